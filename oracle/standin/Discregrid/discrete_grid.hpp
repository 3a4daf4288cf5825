// TEST INFRASTRUCTURE ONLY (oracle/): declaration-level stand-in for Discregrid
// (InteractiveComputerGraphics/Discregrid @ ddf20dc0, reference pin CMake/SetUpExternalProjects.cmake:16-17).
// The reference's TimeStep.h:7 includes this header unconditionally, but the grid is only *called* for the
// Koschier2017/Bender2019 map boundaries (TimeStep.cpp:254-272,400-412), which are out of scope here
// (Akinci2012 only).  The bodies below are inert.
#pragma once
#include <Eigen/Dense>
#include <array>
#include <limits>

namespace Discregrid
{
	class DiscreteGrid
	{
	public:
		virtual ~DiscreteGrid() {}
		bool determineShapeFunctions(unsigned int, Eigen::Vector3d const&, std::array<unsigned int, 32>&, Eigen::Vector3d&,
			Eigen::Matrix<double, 32, 1>&, Eigen::Matrix<double, 32, 3>* = nullptr) const { return false; }
		double interpolate(unsigned int, Eigen::Vector3d const&, Eigen::Vector3d* = nullptr) const { return std::numeric_limits<double>::max(); }
		double interpolate(unsigned int, Eigen::Vector3d const&, std::array<unsigned int, 32> const&, Eigen::Vector3d const&,
			Eigen::Matrix<double, 32, 1> const&, Eigen::Vector3d* = nullptr, Eigen::Matrix<double, 32, 3>* = nullptr) const { return std::numeric_limits<double>::max(); }
	};
}
