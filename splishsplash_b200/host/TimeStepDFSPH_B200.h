// TimeStepDFSPH_B200 -- drop-in replacement for SPH::TimeStepDFSPH (SPlisHSPlasH/DFSPH/TimeStepDFSPH.{h,cpp}) that runs
// the per-step neighbourhood search and the whole DFSPH pressure-solver loop on a B200 through the C ABI of
// include/dfsph_b200.h.  It keeps the reference's plugin surface:
//   * derives SPH::TimeStep (TimeStep.h:13-59): step(), reset(), resize(), init(), getMethodName(), getNumIterations()
//   * the same GenParam parameter names / static handles as TimeStepDFSPH (TimeStepDFSPH.cpp:73-115):
//     iterations, minIterations, maxIterations, maxError, iterationsV, maxIterationsV, maxErrorV, enableDivergenceSolver
//   * the same particle fields on every FluidModel (TimeStepDFSPH.cpp:49-53): "factor", "advected density",
//     "p / rho^2", "p_v / rho^2", "pressure acceleration", served from host mirrors refreshed after each step
//   * reads all state through Simulation::getCurrent() / TimeManager::getCurrent() like the reference (:119-123)
// Host arrays (FluidModel::m_x, m_v, m_density) stay authoritative between steps: every step uploads x and v, and
// downloads x, v and density (dfsph_b200_step_host), so exporters, the GUI and state files keep working unchanged.
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

		// host mirrors of the DFSPH particle fields (SimulationDataDFSPH.h:25-33), fluid model 0
		std::vector<Real> m_factor, m_density_adv, m_pressure_rho2, m_pressure_rho2_V;
		std::vector<Vector3r> m_pressureAccel;
		bool m_syncAllFields;

		void* m_lib;                 // dlopen handle of libdfsph_b200_{f32,f64}.so
		dfsph_b200_ctx* m_ctx;
		bool m_modelUploaded;
		struct Api;
		Api* m_api;

		void loadLibrary(const std::string& path);
		void uploadModel();
		void pushParameters();
		void check(int rc, const char* what);

		virtual void initParameters();

	public:
		static std::string METHOD_NAME;
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

		/** also refresh the five DFSPH field mirrors after every step (default true; exporters/GUI read them) */
		void setSyncAllFields(bool b) { m_syncAllFields = b; }

		/** Neighbour lists of the fluid particles at the CURRENT host positions, in host index space (row i = host
		 *  particle i).  other = 0: fluid neighbours (host indices); other = 1: boundary neighbours, indices into the
		 *  concatenation of all boundary models in the order Simulation holds them.  Lists are ascending.  For tests and
		 *  non-ported host code only -- the solver never materialises host-visible lists. */
		void downloadNeighbors(unsigned int other, std::vector<unsigned int>& offsets, std::vector<unsigned int>& indices);
	};

	/** Facade with the PointSet accessors of CompactNSearch the reference uses (Simulation.h:456-473:
	 *  n_neighbors / neighbor / neighbor_list), served from the device search of a TimeStepDFSPH_B200.
	 *  Point-set numbering follows the reference: 0 = the fluid model, 1.. = the boundary models. */
	class NeighborhoodSearch_B200
	{
	protected:
		TimeStepDFSPH_B200& m_ts;
		std::vector<unsigned int> m_off[2], m_idx[2];
		std::vector<unsigned int> m_bodyStart;                 // first concatenated index of every boundary model
		std::vector<std::vector<unsigned int>> m_scratch;
	public:
		NeighborhoodSearch_B200(TimeStepDFSPH_B200& ts) : m_ts(ts) {}
		/** NeighborhoodSearch::find_neighbors() (Simulation.cpp:617): runs the device search on the host positions. */
		void find_neighbors();
		unsigned int n_neighbors(unsigned int neighborPointSet, unsigned int i) const;
		unsigned int neighbor(unsigned int neighborPointSet, unsigned int i, unsigned int k) const;
		/** neighbour list of fluid particle i inside point set neighborPointSet (local indices of that set) */
		std::vector<unsigned int> neighbor_list(unsigned int neighborPointSet, unsigned int i) const;
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
