// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for the GenericParameters library
// (InteractiveComputerGraphics/GenericParameters @ a4e2744e, pinned in the reference at
// CMake/SetUpExternalProjects.cmake:35-36), which the reference fetches at build time and
// which is absent from /root/reference.  Written from scratch from the reference's call
// sites (TimeStepDFSPH.cpp:73-115, Simulation.cpp:163-277, SceneLoader.cpp:251-346); it
// provides only the API subset those call sites need so that the reference's solver
// sources compile UNMODIFIED into oracle/_ref.  Not part of the product path.
#pragma once
#include <functional>
#include <memory>
#include <string>
#include <vector>
#include <cstring>
#include <cfloat>
#include <climits>

namespace GenParam
{
	class ParameterBase
	{
	public:
		enum DataTypes { BOOL = 0, DOUBLE, ENUM, FLOAT, FUNCTION, INT8, INT16, INT32, LIST, STRING, STRUCT,
			UINT8, UINT16, UINT32, VEC_FLOAT, VEC_DOUBLE, VEC_INT32, VEC_UINT32 };
		template<typename T> using GetFunc = std::function<T()>;
		template<typename T> using SetFunc = std::function<void(T)>;
		template<typename T> using GetVecFunc = std::function<T*()>;
		template<typename T> using SetVecFunc = std::function<void(T*)>;

		ParameterBase(const std::string& name, const std::string& label, DataTypes type)
			: m_name(name), m_label(label), m_type(type), m_readOnly(false), m_visible(true) {}
		virtual ~ParameterBase() {}
		const std::string& getName() const { return m_name; }
		void setName(const std::string& s) { m_name = s; }
		const std::string& getLabel() const { return m_label; }
		void setLabel(const std::string& s) { m_label = s; }
		const std::string& getGroup() const { return m_group; }
		void setGroup(const std::string& s) { m_group = s; }
		const std::string& getDescription() const { return m_description; }
		void setDescription(const std::string& s) { m_description = s; }
		const std::string& getHotKey() const { return m_hotKey; }
		void setHotKey(const std::string& s) { m_hotKey = s; }
		DataTypes getType() const { return m_type; }
		bool getReadOnly() const { return m_readOnly; }
		void setReadOnly(bool b) { m_readOnly = b; }
		bool getVisible() const { return m_visible; }
		void setVisible(bool b) { m_visible = b; }
		virtual bool checkType(ParameterBase::DataTypes t) const { return t == m_type; }
	protected:
		std::string m_name, m_label, m_group, m_description, m_hotKey;
		DataTypes m_type;
		bool m_readOnly, m_visible;
	};

	template<typename T> struct TypeTag;
	template<> struct TypeTag<bool> { static const ParameterBase::DataTypes value = ParameterBase::BOOL; };
	template<> struct TypeTag<double> { static const ParameterBase::DataTypes value = ParameterBase::DOUBLE; };
	template<> struct TypeTag<float> { static const ParameterBase::DataTypes value = ParameterBase::FLOAT; };
	template<> struct TypeTag<char> { static const ParameterBase::DataTypes value = ParameterBase::INT8; };
	template<> struct TypeTag<short> { static const ParameterBase::DataTypes value = ParameterBase::INT16; };
	template<> struct TypeTag<int> { static const ParameterBase::DataTypes value = ParameterBase::INT32; };
	template<> struct TypeTag<unsigned char> { static const ParameterBase::DataTypes value = ParameterBase::UINT8; };
	template<> struct TypeTag<unsigned short> { static const ParameterBase::DataTypes value = ParameterBase::UINT16; };
	template<> struct TypeTag<unsigned int> { static const ParameterBase::DataTypes value = ParameterBase::UINT32; };
	template<> struct TypeTag<std::string> { static const ParameterBase::DataTypes value = ParameterBase::STRING; };
	template<typename T> struct VecTypeTag;
	template<> struct VecTypeTag<float> { static const ParameterBase::DataTypes value = ParameterBase::VEC_FLOAT; };
	template<> struct VecTypeTag<double> { static const ParameterBase::DataTypes value = ParameterBase::VEC_DOUBLE; };
	template<> struct VecTypeTag<int> { static const ParameterBase::DataTypes value = ParameterBase::VEC_INT32; };
	template<> struct VecTypeTag<unsigned int> { static const ParameterBase::DataTypes value = ParameterBase::VEC_UINT32; };

	/** Scalar parameter bound either to a variable or to a getter/setter pair. */
	template<typename T>
	class Parameter : public ParameterBase
	{
	public:
		Parameter(const std::string& name, const std::string& label, DataTypes type, T* valuePtr)
			: ParameterBase(name, label, type)
		{
			m_getValue = [valuePtr]() { return *valuePtr; };
			m_setValue = [valuePtr](T v) { *valuePtr = v; };
		}
		Parameter(const std::string& name, const std::string& label, DataTypes type, GetFunc<T> g, SetFunc<T> s)
			: ParameterBase(name, label, type), m_getValue(g), m_setValue(s) {}
		virtual ~Parameter() {}
		virtual void setValue(const T v) { if (m_setValue != nullptr) m_setValue(v); }
		T getValue() const { return m_getValue(); }
	protected:
		GetFunc<T> m_getValue;
		SetFunc<T> m_setValue;
	};

	template<typename T>
	class NumericParameter : public Parameter<T>
	{
	public:
		NumericParameter(const std::string& name, const std::string& label, T* valuePtr)
			: Parameter<T>(name, label, TypeTag<T>::value, valuePtr), m_hasMin(false), m_hasMax(false) {}
		NumericParameter(const std::string& name, const std::string& label, ParameterBase::GetFunc<T> g, ParameterBase::SetFunc<T> s)
			: Parameter<T>(name, label, TypeTag<T>::value, g, s), m_hasMin(false), m_hasMax(false) {}
		virtual void setValue(const T v)
		{
			T val = v;
			if (m_hasMin && val < m_minValue) val = m_minValue;
			if (m_hasMax && val > m_maxValue) val = m_maxValue;
			Parameter<T>::setValue(val);
		}
		void setMinValue(const T v) { m_minValue = v; m_hasMin = true; }
		void setMaxValue(const T v) { m_maxValue = v; m_hasMax = true; }
		T getMinValue() const { return m_minValue; }
		T getMaxValue() const { return m_maxValue; }
	protected:
		T m_minValue, m_maxValue;
		bool m_hasMin, m_hasMax;
	};

	using FloatParameter = NumericParameter<float>;
	using DoubleParameter = NumericParameter<double>;
	using CharParameter = NumericParameter<char>;
	using ShortParameter = NumericParameter<short>;
	using IntParameter = NumericParameter<int>;
	using UnsignedCharParameter = NumericParameter<unsigned char>;
	using UnsignedShortParameter = NumericParameter<unsigned short>;
	using UnsignedIntParameter = NumericParameter<unsigned int>;

	class BoolParameter : public Parameter<bool>
	{
	public:
		BoolParameter(const std::string& name, const std::string& label, bool* p) : Parameter<bool>(name, label, ParameterBase::BOOL, p) {}
		BoolParameter(const std::string& name, const std::string& label, GetFunc<bool> g, SetFunc<bool> s) : Parameter<bool>(name, label, ParameterBase::BOOL, g, s) {}
	};

	class StringParameter : public Parameter<std::string>
	{
	public:
		StringParameter(const std::string& name, const std::string& label, std::string* p) : Parameter<std::string>(name, label, ParameterBase::STRING, p) {}
		StringParameter(const std::string& name, const std::string& label, GetFunc<std::string> g, SetFunc<std::string> s) : Parameter<std::string>(name, label, ParameterBase::STRING, g, s) {}
	};

	class EnumParameter : public Parameter<int>
	{
	public:
		struct EnumValue { int id; std::string name; };
		EnumParameter(const std::string& name, const std::string& label, int* p) : Parameter<int>(name, label, ParameterBase::ENUM, p), m_idCounter(0) {}
		EnumParameter(const std::string& name, const std::string& label, GetFunc<int> g, SetFunc<int> s) : Parameter<int>(name, label, ParameterBase::ENUM, g, s), m_idCounter(0) {}
		void addEnumValue(const std::string& name, int& id) { id = m_idCounter++; m_enumValues.push_back({ id, name }); }
		const std::vector<EnumValue>& getEnumValues() const { return m_enumValues; }
		void clearEnumValues() { m_enumValues.clear(); m_idCounter = 0; }
	protected:
		int m_idCounter;
		std::vector<EnumValue> m_enumValues;
	};

	template<typename T>
	class VectorParameter : public ParameterBase
	{
	public:
		VectorParameter(const std::string& name, const std::string& label, unsigned int dim, T* valuePtr)
			: ParameterBase(name, label, VecTypeTag<T>::value), m_dim(dim)
		{
			m_getVecValue = [valuePtr]() { return valuePtr; };
			m_setVecValue = [valuePtr, dim](T* v) { std::memcpy(valuePtr, v, dim * sizeof(T)); };
		}
		VectorParameter(const std::string& name, const std::string& label, unsigned int dim, GetVecFunc<T> g, SetVecFunc<T> s)
			: ParameterBase(name, label, VecTypeTag<T>::value), m_dim(dim), m_getVecValue(g), m_setVecValue(s) {}
		void setValue(T* v) { if (m_setVecValue != nullptr) m_setVecValue(v); }
		T* getValue() const { return m_getVecValue(); }
		unsigned int getDim() const { return m_dim; }
	protected:
		unsigned int m_dim;
		GetVecFunc<T> m_getVecValue;
		SetVecFunc<T> m_setVecValue;
	};
	using FloatVectorParameter = VectorParameter<float>;
	using DoubleVectorParameter = VectorParameter<double>;
	using UnsignedIntVectorParameter = VectorParameter<unsigned int>;

	class ParameterObject
	{
	public:
		using ParameterPtr = std::unique_ptr<ParameterBase>;
		ParameterObject() {}
		virtual ~ParameterObject() {}
		virtual void initParameters() {}

		unsigned int numParameters() const { return static_cast<unsigned int>(m_parameters.size()); }
		ParameterBase* getParameter(const unsigned int index) { return m_parameters[index].get(); }
		ParameterBase* const getParameter(const unsigned int index) const { return m_parameters[index].get(); }

		void setVisible(const unsigned int id, const bool v) { getParameter(id)->setVisible(v); }
		void setReadOnly(const unsigned int id, const bool v) { getParameter(id)->setReadOnly(v); }
		void setName(const unsigned int id, const std::string& s) { getParameter(id)->setName(s); }
		void setLabel(const unsigned int id, const std::string& s) { getParameter(id)->setLabel(s); }
		void setGroup(const unsigned int id, const std::string& s) { getParameter(id)->setGroup(s); }
		void setDescription(const unsigned int id, const std::string& s) { getParameter(id)->setDescription(s); }
		void setHotKey(const unsigned int id, const std::string& s) { getParameter(id)->setHotKey(s); }

		template<typename T> int createNumericParameter(const std::string& name, const std::string& label, T* valuePtr)
		{ m_parameters.push_back(ParameterPtr(new NumericParameter<T>(name, label, valuePtr))); return static_cast<int>(m_parameters.size() - 1); }
		template<typename T> int createNumericParameter(const std::string& name, const std::string& label, ParameterBase::GetFunc<T> g, ParameterBase::SetFunc<T> s = {})
		{ m_parameters.push_back(ParameterPtr(new NumericParameter<T>(name, label, g, s))); return static_cast<int>(m_parameters.size() - 1); }
		int createBoolParameter(const std::string& name, const std::string& label, bool* valuePtr)
		{ m_parameters.push_back(ParameterPtr(new BoolParameter(name, label, valuePtr))); return static_cast<int>(m_parameters.size() - 1); }
		int createBoolParameter(const std::string& name, const std::string& label, ParameterBase::GetFunc<bool> g, ParameterBase::SetFunc<bool> s = {})
		{ m_parameters.push_back(ParameterPtr(new BoolParameter(name, label, g, s))); return static_cast<int>(m_parameters.size() - 1); }
		int createEnumParameter(const std::string& name, const std::string& label, int* valuePtr)
		{ m_parameters.push_back(ParameterPtr(new EnumParameter(name, label, valuePtr))); return static_cast<int>(m_parameters.size() - 1); }
		int createEnumParameter(const std::string& name, const std::string& label, ParameterBase::GetFunc<int> g, ParameterBase::SetFunc<int> s = {})
		{ m_parameters.push_back(ParameterPtr(new EnumParameter(name, label, g, s))); return static_cast<int>(m_parameters.size() - 1); }
		int createStringParameter(const std::string& name, const std::string& label, std::string* valuePtr)
		{ m_parameters.push_back(ParameterPtr(new StringParameter(name, label, valuePtr))); return static_cast<int>(m_parameters.size() - 1); }
		int createStringParameter(const std::string& name, const std::string& label, ParameterBase::GetFunc<std::string> g, ParameterBase::SetFunc<std::string> s = {})
		{ m_parameters.push_back(ParameterPtr(new StringParameter(name, label, g, s))); return static_cast<int>(m_parameters.size() - 1); }
		template<typename T> int createVectorParameter(const std::string& name, const std::string& label, const unsigned int dim, T* valuePtr)
		{ m_parameters.push_back(ParameterPtr(new VectorParameter<T>(name, label, dim, valuePtr))); return static_cast<int>(m_parameters.size() - 1); }
		template<typename T> int createVectorParameter(const std::string& name, const std::string& label, const unsigned int dim, ParameterBase::GetVecFunc<T> g, ParameterBase::SetVecFunc<T> s = {})
		{ m_parameters.push_back(ParameterPtr(new VectorParameter<T>(name, label, dim, g, s))); return static_cast<int>(m_parameters.size() - 1); }

		template<typename T> T getValue(const unsigned int id) const
		{ return static_cast<Parameter<T>*>(m_parameters[id].get())->getValue(); }
		template<typename T> void setValue(const unsigned int id, const T v)
		{ static_cast<Parameter<T>*>(m_parameters[id].get())->setValue(v); }
		template<typename T> T* getVecValue(const unsigned int id) const
		{ return static_cast<VectorParameter<T>*>(m_parameters[id].get())->getValue(); }
		template<typename T> void setVecValue(const unsigned int id, T* v)
		{ static_cast<VectorParameter<T>*>(m_parameters[id].get())->setValue(v); }

	protected:
		std::vector<ParameterPtr> m_parameters;
	};
}
