// TimeStepDFSPH_B200 -- drop-in replacement for SPH::TimeStepDFSPH (SPlisHSPlasH/DFSPH/TimeStepDFSPH.{h,cpp}) that runs
// the per-step neighbourhood search and the whole DFSPH pressure-solver loop on a B200 through the C ABI of
// include/dfsph_b200.h.  It keeps the reference's plugin surface:
//   * derives SPH::TimeStep (TimeStep.h:13-59): step(), reset(), resize(), init(), getMethodName(), getNumIterations()
//   * the same GenParam parameter names / static handles as TimeStepDFSPH (TimeStepDFSPH.cpp:73-115):
//     iterations, minIterations, maxIterations, maxError, iterationsV, maxIterationsV, maxErrorV, enableDivergenceSolver
//   * the same particle fields on every FluidModel (TimeStepDFSPH.cpp:49-53): "factor", "advected density",
//     "p / rho^2", "p_v / rho^2", "pressure acceleration", served from host mirrors that are refreshed lazily (the first
//     reader of a field after a step triggers one bulk download of that field)
//   * reads all state through Simulation::getCurrent() / TimeManager::getCurrent() like the reference (:119-123)
// Host arrays (FluidModel::m_x, m_v, m_density) stay authoritative between steps by default: every step uploads x and v
// and downloads x, v and density (dfsph_b200_step_host), so exporters, the GUI and state files keep working unchanged;
// setHostStateSync(false) keeps the state on the device and downloadState() fetches it on demand.
//
// The CUDA library is loaded with dlopen() (libdfsph_b200_f32.so for float builds of the reference,
// libdfsph_b200_f64.so for USE_DOUBLE builds); there is no CPU fallback: if the library or a CUDA device is missing
// the constructor throws std::runtime_error.
//
// Scope (checked in resize(), clear error instead of silently different physics): one fluid model, static
// Akinci2012 boundary models, 3-D, no emitters / non-pressure forces (viscosityMethod 0 etc.).
#pragma once

#include "SPlisHSPlasH/Common.h"
#include "SPlisHSPlasH/TimeStep.h"
#include "SPlisHSPlasH/Simulation.h"
#include <cstddef>
#include <string>
#include <vector>

struct dfsph_b200_ctx;

namespace SPH
{
	class TimeStepDFSPH_B200 : public TimeStep
	{
	protected:
		unsigned int m_iterations;
		Real m_maxError;
		unsigned int m_minIterations;
		unsigned int m_maxIterations;
		bool m_enableDivergenceSolver;
		unsigned int m_iterationsV;
		Real m_maxErrorV;
		unsigned int m_maxIterationsV;

		// host mirrors of the DFSPH particle fields (SimulationDataDFSPH.h:25-33), fluid model 0.  LAZY: a step only marks
		// them stale; the first getFct(i) of a field after a step downloads that field once, in bulk (by particle id), so
		// the reference's per-particle readers (SimulatorBase.cpp:2094-2123, exporters, GUI colouring) cost one D2H per
		// field they actually touch and a run without readers costs none.
		enum { F_FACTOR = 0, F_DENSITY_ADV, F_KAPPA, F_KAPPA_V, F_PRESSURE_ACCEL, NUM_MIRRORS };
		std::vector<Real> m_factor, m_density_adv, m_pressure_rho2, m_pressure_rho2_V;
		std::vector<Vector3r> m_pressureAccel;
		bool m_stale[NUM_MIRRORS];
		unsigned int m_fieldDownloads;       // bulk downloads issued so far (tests)
		bool m_syncAllFields;                // eager mode: refresh all five after every step
		bool m_hostStateSync;                // true: FluidModel x, v, density are uploaded / downloaded every step
		bool m_hostStateStale;               // device-resident mode: host x, v, density are behind the device
		unsigned int m_uploadedParticles;

		void* m_lib;                 // dlopen handle of libdfsph_b200_{f32,f64}.so
		dfsph_b200_ctx* m_ctx;
		bool m_modelUploaded;
		struct Api;
		Api* m_api;

		unsigned char m_lastParams[256];     // the dfsph_b200_params pushed last (set_params only when something changed)
		bool m_paramsPushed;

		void loadLibrary(const std::string& path);
		void uploadModel();
		void pushParameters();
		void check(int rc, const char* what);
		void* mirror(int which, unsigned int i);
		std::vector<void*> m_pinned;         // FluidModel arrays page-locked by uploadModel()
		void unpinHostArrays();

		virtual void initParameters();

	public:
		static std::string METHOD_NAME;
		/** The scene loader reads the solver's parameters from Configuration["<getMethodName()>"] (SceneLoader.cpp:231-249).
		 *  Shipped scenes carry a "DFSPH" block with exactly this class's parameter names: call this before the scene is
		 *  read and they run unmodified (the method then reports itself as "DFSPH" everywhere, like the solver it replaces).
		 *  patches/scene_loader_dfsph_block_fallback.patch is the alternative that keeps the name "DFSPH_B200". */
		static void useReferenceConfigBlock() { METHOD_NAME = "DFSPH"; }
		static int SOLVER_ITERATIONS;
		static int MIN_ITERATIONS;
		static int MAX_ITERATIONS;
		static int MAX_ERROR;
		static int SOLVER_ITERATIONS_V;
		static int MAX_ITERATIONS_V;
		static int MAX_ERROR_V;
		static int USE_DIVERGENCE_SOLVER;

		/** libraryPath: directory holding libdfsph_b200_*.so (default: $DFSPH_B200_LIB_DIR or the loader path). */
		TimeStepDFSPH_B200(const std::string& libraryPath = "");
		virtual ~TimeStepDFSPH_B200(void);

		virtual void step();
		virtual void reset();
		virtual void resize();
		virtual std::string getMethodName() { return METHOD_NAME; }
		virtual int getNumIterations() { return m_iterations; }
		/** the device re-sorts the particles into z-order every step; host arrays keep their order */
		virtual void performNeighborhoodSearchSort() {}

		virtual void emittedParticles(FluidModel* model, const unsigned int startIndex);

		/** eager mode: refresh the five DFSPH field mirrors after every step (default false: they are fetched on first use) */
		void setSyncAllFields(bool b) { m_syncAllFields = b; }
		/** number of bulk field downloads so far (one per field touched after a step) */
		unsigned int numFieldDownloads() const { return m_fieldDownloads; }
		/** Bulk accessor for state files / exporters: the whole field in host particle order with ONE
		 *  device-to-host copy, instead of getFct(i) per particle.  name = a FieldDescription name of this method or
		 *  "position" / "velocity" / "density"; dst holds numActiveParticles() * (1 or 3) Reals. */
		void downloadField(const std::string& name, Real* dst);
		/** false: keep x, v on the device between steps (no per-step H2D/D2H of the FluidModel arrays); call
		 *  downloadState() before host code reads FluidModel positions / velocities / densities.  Default true. */
		void setHostStateSync(bool b);
		/** refresh FluidModel x, v, density from the device (no-op when they are current) */
		void downloadState();

		/** Neighbour lists of the fluid particles at the CURRENT host positions, in host index space (row i = host
		 *  particle i).  other = 0: fluid neighbours (host indices); other = 1: boundary neighbours, indices into the
		 *  concatenation of all boundary models in the order Simulation holds them.  Lists are ascending.  For tests and
		 *  non-ported host code only -- the solver never materialises host-visible lists. */
		void downloadNeighbors(unsigned int other, std::vector<unsigned int>& offsets, std::vector<unsigned int>& indices);
		/** the device's current particle order: table[r] = host index of the particle in device row r (NeighborhoodSearch_B200::z_sort) */
		void downloadSortTable(std::vector<unsigned int>& table);
	};

	/** Facade with the part of CompactNSearch::NeighborhoodSearch the reference calls, served from the
	 *  device search of a TimeStepDFSPH_B200.  Point-set numbering follows the reference: 0 = the fluid model,
	 *  1.. = the boundary models (the order Simulation holds them).  The solver itself never goes through this class --
	 *  it exists for tests and for host code of the reference that still wants neighbour lists
	 *  (Simulation.h:451-473, FluidModel.cpp:214-216, 322-346, BoundaryModel_Akinci2012.cpp:108-126, Simulation.cpp:617-626,
	 *  698-755).  Lists are held per (fluid, set) pair in CSR form, ascending, so every accessor is O(1). */
	class NeighborhoodSearch_B200
	{
	public:
		/** CompactNSearch::PointSet view */
		class PointSet
		{
			friend class NeighborhoodSearch_B200;
			NeighborhoodSearch_B200* m_ns; unsigned int m_index;
			Real const* m_x; std::size_t m_n; bool m_dynamic, m_search, m_find; void* m_user;
		public:
			std::size_t n_points() const { return m_n; }
			bool is_dynamic() const { return m_dynamic; }
			void* get_user_data() const { return m_user; }
			unsigned int n_neighbors(unsigned int set, unsigned int i) const { return m_ns->n_neighbors_of(m_index, set, i); }
			unsigned int neighbor(unsigned int set, unsigned int i, unsigned int k) const { return m_ns->neighbor_of(m_index, set, i, k); }
			const std::vector<unsigned int>& neighbor_list(unsigned int set, unsigned int i) const { return m_ns->list_of(m_index, set, i); }
			/** PointSet::sort_field (FluidModel.cpp:339-346): permute `field` by the permutation of the last z_sort() */
			template <typename T> void sort_field(T* field) const
			{
				const std::vector<unsigned int>& t = m_ns->m_sortTable[m_index];
				if (t.empty()) return;
				std::vector<T> tmp(field, field + t.size());
				for (std::size_t i = 0; i < t.size(); i++) field[i] = tmp[t[i]];
			}
		};
	protected:
		TimeStepDFSPH_B200& m_ts;
		std::vector<PointSet> m_sets;
		std::vector<std::vector<bool>> m_active;                 // activation table (set_active)
		// neighbours of the fluid set: CSR per neighbour set (index 0 = fluid, b + 1 = boundary model b), local indices
		std::vector<std::vector<unsigned int>> m_off, m_idx;
		std::vector<std::vector<unsigned int>> m_sortTable;
		mutable std::vector<unsigned int> m_scratch;
		static const std::vector<unsigned int> s_empty;
		unsigned int n_neighbors_of(unsigned int a, unsigned int set, unsigned int i) const;
		unsigned int neighbor_of(unsigned int a, unsigned int set, unsigned int i, unsigned int k) const;
		const std::vector<unsigned int>& list_of(unsigned int a, unsigned int set, unsigned int i) const;
	public:
		NeighborhoodSearch_B200(TimeStepDFSPH_B200& ts) : m_ts(ts) {}
		/** add_point_set (FluidModel.cpp:323: fluid, all flags true; BoundaryModel_Akinci2012.cpp:109: found, does not search).
		 *  Sets must be added in the reference's order: the fluid model first, then the boundary models. */
		unsigned int add_point_set(Real const* x, std::size_t n, bool is_dynamic = true, bool search_neighbors = true, bool find_neighbors = true, void* user_data = nullptr);
		void resize_point_set(unsigned int i, Real const* x, std::size_t n);
		void update_point_sets() {}
		void set_radius(Real) {}                                  // the radius is the solver's support radius (Simulation.cpp:283)
		/** set_active overloads of Simulation.cpp:713-754 */
		void set_active(bool active);
		void set_active(unsigned int i, bool search_neighbors, bool find_neighbors);
		void set_active(unsigned int i, unsigned int j, bool active);
		bool is_active(unsigned int i, unsigned int j) const { return i < m_active.size() && j < m_active[i].size() && m_active[i][j]; }
		std::size_t n_point_sets() const { return m_sets.size(); }
		const PointSet& point_set(unsigned int i) const { return m_sets[i]; }
		const std::vector<PointSet>& point_sets() const { return m_sets; }
		/** NeighborhoodSearch::find_neighbors() (Simulation.cpp:617): runs the device search on the host positions. */
		void find_neighbors();
		/** NeighborhoodSearch::z_sort() (Simulation.cpp:626): fetches the device's particle order as the sort table of set 0 */
		void z_sort();
		// shorthand for set 0 (the fluid), kept from round 1
		unsigned int n_neighbors(unsigned int neighborPointSet, unsigned int i) const { return n_neighbors_of(0, neighborPointSet, i); }
		unsigned int neighbor(unsigned int neighborPointSet, unsigned int i, unsigned int k) const { return neighbor_of(0, neighborPointSet, i, k); }
		/** neighbour list of fluid particle i inside point set neighborPointSet (local indices of that set) */
		std::vector<unsigned int> neighbor_list(unsigned int neighborPointSet, unsigned int i) const { return list_of(0, neighborPointSet, i); }
	};

	/** Registers the method without editing Simulation.cpp: a Simulation subclass that installs TimeStepDFSPH_B200.
	 *  Usage:  Simulation::setCurrent(new Simulation_B200());  ... build the scene as usual ...
	 *          static_cast<Simulation_B200*>(Simulation::getCurrent())->useB200Solver();
	 *  (doc/creating_pressure.md:137-162 describes the alternative: add an enum value + else-if branch to
	 *  Simulation::setSimulationMethod -- the patch is shown in INTEGRATION.md.) */
	class Simulation_B200 : public Simulation
	{
	public:
		void useB200Solver(const std::string& libraryPath = "");
	};
}
