// TEST INFRASTRUCTURE ONLY (oracle/): C-ABI harness around the reference's OWN solver classes, compiled from
// /root/reference UNMODIFIED (see oracle/Makefile).  It builds a scene directly through Simulation::addFluidModel /
// BoundaryModel_Akinci2012::initModel in the order SimulatorBase does (Simulator/SimulatorBase.cpp:432-577:
// Simulation::init -> fluid models -> parameters/simulationMethod -> boundary models -> setSimulationInitialized ->
// StaticBoundarySimulator::deferredInit (StaticBoundarySimulator.cpp:181-199)), bypassing SimulatorBase, the scene
// loader, PBD, GUI and Partio, none of which are on the hot path.  One Real per shared object
// (libsplish_ref_f32.so: float + AVX variant, libsplish_ref_f64.so: double + scalar variant).
// Used by tests/ (parity checker) and by bench.py's cpu_baseline / --impl reference legs only.
#include "SPlisHSPlasH/Common.h"
#include "SPlisHSPlasH/Simulation.h"
#include "SPlisHSPlasH/TimeManager.h"
#include "SPlisHSPlasH/TimeStep.h"
#include "SPlisHSPlasH/FluidModel.h"
#include "SPlisHSPlasH/BoundaryModel_Akinci2012.h"
#include "SPlisHSPlasH/StaticRigidBody.h"
#include "SPlisHSPlasH/DFSPH/TimeStepDFSPH.h"
#include "SPlisHSPlasH/Viscosity/Viscosity_Standard.h"
#include "TimeStepDFSPH_B200.h"   // the product's drop-in solver (splishsplash_b200/host), exercised through the reference stack
#include "Utilities/Timing.h"
#include "Utilities/Counting.h"
#include "Utilities/Logger.h"
#include <chrono>
#include <cstring>
#include <omp.h>

INIT_LOGGING
INIT_TIMING
INIT_COUNTING

using namespace SPH;

namespace {
	double g_step_seconds = 0.0;
	std::string g_err;
	bool g_b200 = false;

	const Real* fieldPtr(FluidModel* fm, const char* name, unsigned int i)
	{
		const FieldDescription& f = fm->getField(name);
		return (const Real*)f.getFct(i);
	}
}

extern "C" {

int ref_sizeof_real() { return (int)sizeof(Real); }
int ref_uses_avx()
{
#ifdef USE_AVX
	return 1;
#else
	return 0;
#endif
}
int ref_num_threads() { return omp_get_max_threads(); }
void ref_set_num_threads(int n) { omp_set_num_threads(n); }

/* Simulation::init (Simulation.cpp:131-152). */
int ref_create(double particleRadius)
{
	if (Simulation::hasCurrent()) return -1;
	Utilities::Timing::m_dontPrintTimes = true;
	Simulation::setCurrent(new Simulation_B200());   // plain Simulation behaviour unless ref_configure_b200 is used
	Simulation* sim = Simulation::getCurrent();
	sim->init(static_cast<Real>(particleRadius), false);
	g_b200 = false;
	return 0;
}

/* Simulation::addFluidModel (Simulation.cpp:684-689); nMaxEmitterParticles = 0. */
int ref_add_fluid(const Real* x, const Real* v, unsigned int n, double density0)
{
	Simulation* sim = Simulation::getCurrent();
	std::vector<Vector3r> xs(n), vs(n);
	std::vector<unsigned int> ids(n, 0u);
	for (unsigned int i = 0; i < n; i++)
	{
		xs[i] = Vector3r(x[3 * i], x[3 * i + 1], x[3 * i + 2]);
		vs[i] = v ? Vector3r(v[3 * i], v[3 * i + 1], v[3 * i + 2]) : Vector3r::Zero();
	}
	const unsigned int idx = sim->numberOfFluidModels();
	sim->addFluidModel("Fluid" + std::to_string(idx), n, xs.data(), vs.data(), ids.data(), 0);
	FluidModel* fm = sim->getFluidModel(idx);
	fm->setDensity0(static_cast<Real>(density0));
	fm->setViscosityMethod(0u);   // no non-pressure forces
	return (int)idx;
}

/* Configuration step: boundaryHandlingMethod Akinci2012, simulationMethod DFSPH (creates TimeStepDFSPH,
   Simulation.cpp:577-583), then the kernel choice is re-applied AFTER the method (SURVEY.md 3.4 ordering trap).
   kernel: 0 = cubic, 4 = precomputed cubic (the DFSPH default). */
int ref_configure(int kernel, int gradKernel)
{
	Simulation* sim = Simulation::getCurrent();
	sim->setValue<int>(Simulation::BOUNDARY_HANDLING_METHOD, Simulation::ENUM_AKINCI2012);
	sim->setValue<int>(Simulation::SIMULATION_METHOD, Simulation::ENUM_SIMULATION_DFSPH);
	sim->setValue<int>(Simulation::KERNEL_METHOD, kernel);
	sim->setValue<int>(Simulation::GRAD_KERNEL_METHOD, gradKernel);
	return 0;
}

/* Same configuration step, but the solver is the B200 drop-in (TimeStepDFSPH_B200 through Simulation_B200): the
   reference's Simulation, FluidModel, BoundaryModel_Akinci2012 and TimeManager objects stay in charge of everything
   else.  libdir: directory holding libdfsph_b200_*.so.  Returns -2 if the CUDA library / device is unavailable. */
int ref_configure_b200(int kernel, int gradKernel, const char* libdir)
{
	Simulation* sim = Simulation::getCurrent();
	sim->setValue<int>(Simulation::BOUNDARY_HANDLING_METHOD, Simulation::ENUM_AKINCI2012);
	try {
		static_cast<Simulation_B200*>(sim)->useB200Solver(libdir ? libdir : "");
	} catch (const std::exception& e) {
		g_err = e.what();
		return -2;
	}
	sim->setValue<int>(Simulation::KERNEL_METHOD, kernel);
	sim->setValue<int>(Simulation::GRAD_KERNEL_METHOD, gradKernel);
	g_b200 = true;
	return 0;
}
/* Option B of INTEGRATION.md: the solver is selected like a built-in method, through the "simulationMethod" enum
   parameter.  Only a library whose Simulation.{h,cpp} carry patches/register_dfsph_b200.patch knows the id (7); the
   unpatched reference falls back to DFSPH for an unknown id (Simulation.cpp:538-539), reported as -3 here.
   TimeStepDFSPH_B200's default constructor finds the CUDA library through DFSPH_B200_LIB_DIR. */
int ref_configure_by_method_id(int method, int kernel, int gradKernel)
{
	Simulation* sim = Simulation::getCurrent();
	sim->setValue<int>(Simulation::BOUNDARY_HANDLING_METHOD, Simulation::ENUM_AKINCI2012);
	try {
		sim->setValue<int>(Simulation::SIMULATION_METHOD, method);
	} catch (const std::exception& e) {
		g_err = e.what();
		return -2;
	}
	if (dynamic_cast<TimeStepDFSPH_B200*>(sim->getTimeStep()) == nullptr) {
		g_err = "simulationMethod " + std::to_string(method) + " is not DFSPH_B200 in this build (got " + sim->getTimeStep()->getMethodName() + ")";
		return -3;
	}
	sim->setValue<int>(Simulation::KERNEL_METHOD, kernel);
	sim->setValue<int>(Simulation::GRAD_KERNEL_METHOD, gradKernel);
	g_b200 = true;
	return 0;
}
const char* ref_last_error() { return g_err.c_str(); }

/* Static Akinci2012 boundary (StaticBoundarySimulator.cpp:146-152 + BoundaryModel_Akinci2012::initModel). */
int ref_add_boundary(const Real* x, unsigned int n)
{
	Simulation* sim = Simulation::getCurrent();
	std::vector<Vector3r> xs(n);
	for (unsigned int i = 0; i < n; i++) xs[i] = Vector3r(x[3 * i], x[3 * i + 1], x[3 * i + 2]);
	StaticRigidBody* rb = new StaticRigidBody();
	rb->setPosition0(Vector3r::Zero());
	rb->setPosition(Vector3r::Zero());
	rb->setRotation0(Quaternionr::Identity());
	rb->setRotation(Quaternionr::Identity());
	BoundaryModel_Akinci2012* bm = new BoundaryModel_Akinci2012();
	bm->initModel(rb, n, xs.data());
	sim->addBoundaryModel(bm);
	return (int)sim->numberOfBoundaryModels() - 1;
}

/* SimulatorBase::deferredInit tail (SimulatorBase.cpp:575-576) + StaticBoundarySimulator::deferredInit. */
int ref_finalize()
{
	Simulation* sim = Simulation::getCurrent();
	sim->setSimulationInitialized(true);
	sim->performNeighborhoodSearchSort();
	sim->updateBoundaryVolume();
	return 0;
}

int ref_set_real(const char* name, double v)
{
	Simulation* sim = Simulation::getCurrent();
	TimeStep* ts = sim->getTimeStep();
	const std::string s(name);
	if (s == "timeStepSize") TimeManager::getCurrent()->setTimeStepSize(static_cast<Real>(v));
	else if (s == "maxError") ts->setValue<Real>(g_b200 ? TimeStepDFSPH_B200::MAX_ERROR : TimeStepDFSPH::MAX_ERROR, static_cast<Real>(v));
	else if (s == "maxErrorV") ts->setValue<Real>(g_b200 ? TimeStepDFSPH_B200::MAX_ERROR_V : TimeStepDFSPH::MAX_ERROR_V, static_cast<Real>(v));
	else if (s == "cflFactor") sim->setValue<Real>(Simulation::CFL_FACTOR, static_cast<Real>(v));
	else if (s == "cflMinTimeStepSize") sim->setValue<Real>(Simulation::CFL_MIN_TIMESTEPSIZE, static_cast<Real>(v));
	else if (s == "cflMaxTimeStepSize") sim->setValue<Real>(Simulation::CFL_MAX_TIMESTEPSIZE, static_cast<Real>(v));
	else return -1;
	return 0;
}

int ref_set_int(const char* name, int v)
{
	Simulation* sim = Simulation::getCurrent();
	TimeStep* ts = sim->getTimeStep();
	const std::string s(name);
	if (s == "minIterations") ts->setValue<unsigned int>(g_b200 ? TimeStepDFSPH_B200::MIN_ITERATIONS : TimeStepDFSPH::MIN_ITERATIONS, (unsigned int)v);
	else if (s == "maxIterations") ts->setValue<unsigned int>(g_b200 ? TimeStepDFSPH_B200::MAX_ITERATIONS : TimeStepDFSPH::MAX_ITERATIONS, (unsigned int)v);
	else if (s == "maxIterationsV") ts->setValue<unsigned int>(g_b200 ? TimeStepDFSPH_B200::MAX_ITERATIONS_V : TimeStepDFSPH::MAX_ITERATIONS_V, (unsigned int)v);
	else if (s == "enableDivergenceSolver") ts->setValue<bool>(g_b200 ? TimeStepDFSPH_B200::USE_DIVERGENCE_SOLVER : TimeStepDFSPH::USE_DIVERGENCE_SOLVER, v != 0);
	else if (s == "cflMethod") sim->setValue<int>(Simulation::CFL_METHOD, v);
	else if (s == "enableZSort") sim->setValue<bool>(Simulation::ENABLE_Z_SORT, v != 0);
	else if (s == "stepsPerZSort") sim->setValue<unsigned int>(Simulation::STEPS_PER_Z_SORT, (unsigned int)v);
	else return -1;
	return 0;
}

/* FluidModel "viscosityMethod" (0 none, 1 "Standard viscosity") + Viscosity_Standard parameters. */
int ref_set_viscosity(int fluid, int method, double viscosity, double viscosityBoundary)
{
	FluidModel* fm = Simulation::getCurrent()->getFluidModel(fluid);
	fm->setViscosityMethod((unsigned int)method);
	if (method == 1)
	{
		NonPressureForceBase* v = fm->getViscosityBase();
		v->setValue<Real>(Viscosity_Standard::VISCOSITY_COEFFICIENT, static_cast<Real>(viscosity));
		v->setValue<Real>(Viscosity_Standard::VISCOSITY_COEFFICIENT_BOUNDARY, static_cast<Real>(viscosityBoundary));
	}
	return 0;
}

int ref_set_gravity(double gx, double gy, double gz)
{
	Real g[3] = { static_cast<Real>(gx), static_cast<Real>(gy), static_cast<Real>(gz) };
	Simulation::getCurrent()->setVecValue<Real>(Simulation::GRAVITATION, g);
	return 0;
}

/* TimeStepDFSPH::step() (TimeStepDFSPH.cpp:117-249), n times; wall time accumulated. */
int ref_step(int n)
{
	Simulation* sim = Simulation::getCurrent();
	for (int k = 0; k < n; k++)
	{
		auto t0 = std::chrono::high_resolution_clock::now();
		sim->getTimeStep()->step();
		auto t1 = std::chrono::high_resolution_clock::now();
		g_step_seconds += std::chrono::duration<double>(t1 - t0).count();
	}
	return 0;
}

double ref_step_seconds() { return g_step_seconds; }
void ref_reset_step_seconds() { g_step_seconds = 0.0; }

/* Average of one of the reference's own START_TIMING timers (Utilities/Timing.h), in ms; -1 if unknown. */
double ref_avg_timer_ms(const char* name)
{
	for (auto& kv : Utilities::Timing::m_averageTimes)
		if (kv.second.name == name && kv.second.counter > 0) return kv.second.totalTime / kv.second.counter;
	return -1.0;
}

/* Neighbour search + density only (what ReadWriteStateTests.cpp:349-350 does). */
int ref_search_and_density()
{
	Simulation* sim = Simulation::getCurrent();
	sim->performNeighborhoodSearch();
	for (unsigned int m = 0; m < sim->numberOfFluidModels(); m++) sim->getTimeStep()->computeDensities(m);
	return 0;
}

unsigned int ref_num_particles(int fluid) { return Simulation::getCurrent()->getFluidModel(fluid)->numActiveParticles(); }
unsigned int ref_num_boundary_particles(int b) { return static_cast<BoundaryModel_Akinci2012*>(Simulation::getCurrent()->getBoundaryModel(b))->numberOfParticles(); }
double ref_time() { return TimeManager::getCurrent()->getTime(); }
double ref_time_step_size() { return TimeManager::getCurrent()->getTimeStepSize(); }
int ref_iterations() { return Simulation::getCurrent()->getTimeStep()->getValue<unsigned int>(g_b200 ? TimeStepDFSPH_B200::SOLVER_ITERATIONS : TimeStepDFSPH::SOLVER_ITERATIONS); }
int ref_iterations_v() { return Simulation::getCurrent()->getTimeStep()->getValue<unsigned int>(g_b200 ? TimeStepDFSPH_B200::SOLVER_ITERATIONS_V : TimeStepDFSPH::SOLVER_ITERATIONS_V); }
const char* ref_method_name() { static std::string s; s = Simulation::getCurrent()->getTimeStep()->getMethodName(); return s.c_str(); }
double ref_w_zero() { return Simulation::getCurrent()->W_zero(); }
double ref_fluid_volume(int fluid) { return Simulation::getCurrent()->getFluidModel(fluid)->getVolume(0); }
int ref_kernel() { return Simulation::getCurrent()->getKernel(); }

/* The reference's kernel classes evaluated pointwise (SPHKernels.h): kind = Simulation "kernel" id 0..4.  Lets the
   restatement's kernel functions be pinned against the reference's directly (tests/test_oracle.py). */
int ref_eval_kernel(int kind, unsigned int n, const Real* r, Real* W, Real* gradW)
{
	for (unsigned int i = 0; i < n; i++)
	{
		const Vector3r x(r[3 * i], r[3 * i + 1], r[3 * i + 2]);
		Real w; Vector3r g;
		switch (kind)
		{
			case 0: w = CubicKernel::W(x); g = CubicKernel::gradW(x); break;
			case 1: w = WendlandQuinticC2Kernel::W(x); g = WendlandQuinticC2Kernel::gradW(x); break;
			case 2: w = Poly6Kernel::W(x); g = Poly6Kernel::gradW(x); break;
			case 3: w = SpikyKernel::W(x); g = SpikyKernel::gradW(x); break;
			case 4: w = Simulation::PrecomputedCubicKernel::W(x); g = Simulation::PrecomputedCubicKernel::gradW(x); break;
			default: return -1;
		}
		W[i] = w; gradW[3 * i] = g[0]; gradW[3 * i + 1] = g[1]; gradW[3 * i + 2] = g[2];
	}
	return 0;
}

/* Field names are the reference's own FieldDescription names (FluidModel.cpp:60-66, TimeStepDFSPH.cpp:49-53):
   "position", "velocity", "density", "factor", "advected density", "p / rho^2", "p_v / rho^2",
   "pressure acceleration"; plus "acceleration".  dim = 1 or 3.  Output in current (z-sorted) array order. */
int ref_get_field(int fluid, const char* name, Real* out, int dim)
{
	FluidModel* fm = Simulation::getCurrent()->getFluidModel(fluid);
	const unsigned int n = fm->numActiveParticles();
	if (std::string(name) == "acceleration")
	{
		for (unsigned int i = 0; i < n; i++) for (int k = 0; k < 3; k++) out[3 * i + k] = fm->getAcceleration(i)[k];
		return 0;
	}
	for (unsigned int i = 0; i < n; i++)
	{
		const Real* p = fieldPtr(fm, name, i);
		for (int k = 0; k < dim; k++) out[dim * i + k] = p[k];
	}
	return 0;
}

int ref_get_ids(int fluid, unsigned int* out)
{
	FluidModel* fm = Simulation::getCurrent()->getFluidModel(fluid);
	for (unsigned int i = 0; i < fm->numActiveParticles(); i++) out[i] = fm->getParticleId(i);
	return 0;
}

/* Overwrite state (used to start reference and device runs from an identical, non-trivial state). */
int ref_set_state(int fluid, const Real* x, const Real* v)
{
	FluidModel* fm = Simulation::getCurrent()->getFluidModel(fluid);
	for (unsigned int i = 0; i < fm->numActiveParticles(); i++)
	{
		if (x) fm->getPosition(i) = Vector3r(x[3 * i], x[3 * i + 1], x[3 * i + 2]);
		if (v) fm->getVelocity(i) = Vector3r(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
	}
	return 0;
}

/* Overwrite a scalar/vector particle field (array order) through the reference's own FieldDescription accessor. */
int ref_set_field(int fluid, const char* name, const Real* in, int dim)
{
	FluidModel* fm = Simulation::getCurrent()->getFluidModel(fluid);
	const unsigned int n = fm->numActiveParticles();
	const FieldDescription& f = fm->getField(name);
	for (unsigned int i = 0; i < n; i++)
	{
		Real* p = (Real*)f.getFct(i);
		for (int k = 0; k < dim; k++) p[k] = in[dim * i + k];
	}
	return 0;
}

int ref_get_boundary(int b, Real* x, Real* V)
{
	BoundaryModel_Akinci2012* bm = static_cast<BoundaryModel_Akinci2012*>(Simulation::getCurrent()->getBoundaryModel(b));
	for (unsigned int i = 0; i < bm->numberOfParticles(); i++)
	{
		if (x) for (int k = 0; k < 3; k++) x[3 * i + k] = bm->getPosition(i)[k];
		if (V) V[i] = bm->getVolume(i);
	}
	return 0;
}

/* Neighbour counts of fluid set `fluid` against point set `pid` (valid after a step / search). */
int ref_neighbor_counts(int fluid, int pid, unsigned int* counts)
{
	Simulation* sim = Simulation::getCurrent();
	const unsigned int n = sim->getFluidModel(fluid)->numActiveParticles();
	for (unsigned int i = 0; i < n; i++) counts[i] = sim->numberOfNeighbors(fluid, pid, i);
	return 0;
}

/* CSR fill; offsets has n+1 entries (caller computes from counts), idx receives neighbour indices (array order). */
int ref_neighbor_lists(int fluid, int pid, const unsigned long long* offsets, unsigned int* idx)
{
	Simulation* sim = Simulation::getCurrent();
	const unsigned int n = sim->getFluidModel(fluid)->numActiveParticles();
	#pragma omp parallel for schedule(static)
	for (int i = 0; i < (int)n; i++)
	{
		const unsigned int c = sim->numberOfNeighbors(fluid, pid, i);
		for (unsigned int k = 0; k < c; k++) idx[offsets[i] + k] = sim->getNeighbor(fluid, pid, i, k);
	}
	return 0;
}

/* Neighbour counts through the drop-in's CompactNSearch-style facade (NeighborhoodSearch_B200) for point set `pid`
   (0 fluid, 1.. boundary models), host index order.  Only valid after ref_configure_b200. */
int ref_b200_neighbor_counts(int pid, unsigned int* counts, unsigned long long* checksum)
{
	if (!g_b200) return -1;
	Simulation* sim = Simulation::getCurrent();
	TimeStepDFSPH_B200* ts = static_cast<TimeStepDFSPH_B200*>(sim->getTimeStep());
	try {
		NeighborhoodSearch_B200 ns(*ts);
		ns.find_neighbors();
		const unsigned int n = sim->getFluidModel(0)->numActiveParticles();
		unsigned long long cs = 0;
		for (unsigned int i = 0; i < n; i++)
		{
			counts[i] = ns.n_neighbors((unsigned int)pid, i);
			for (unsigned int k = 0; k < counts[i]; k++) cs += (unsigned long long)(i + 1) * (ns.neighbor((unsigned int)pid, i, k) + 7u);
		}
		*checksum = cs;
	} catch (const std::exception& e) { g_err = e.what(); return -2; }
	return 0;
}

/* the same checksum over the reference's own lists (array order = host order while no z-sort happened) */
int ref_neighbor_checksum(int fluid, int pid, unsigned long long* checksum)
{
	Simulation* sim = Simulation::getCurrent();
	const unsigned int n = sim->getFluidModel(fluid)->numActiveParticles();
	unsigned long long cs = 0;
	for (unsigned int i = 0; i < n; i++)
		for (unsigned int k = 0; k < sim->numberOfNeighbors(fluid, pid, i); k++)
			cs += (unsigned long long)(i + 1) * (sim->getNeighbor(fluid, pid, i, k) + 7u);
	*checksum = cs;
	return 0;
}

/* ---- drop-in features (valid after ref_configure_b200): lazy field mirrors, bulk field accessor, device-resident state,
   CompactNSearch facade ---- */
static TimeStepDFSPH_B200* b200_ts()
{
	if (!g_b200) return nullptr;
	return dynamic_cast<TimeStepDFSPH_B200*>(Simulation::getCurrent()->getTimeStep());
}
int ref_b200_field_downloads() { TimeStepDFSPH_B200* ts = b200_ts(); return ts ? (int)ts->numFieldDownloads() : -1; }
int ref_b200_set_host_sync(int on)
{
	TimeStepDFSPH_B200* ts = b200_ts();
	if (!ts) return -1;
	try { ts->setHostStateSync(on != 0); } catch (const std::exception& e) { g_err = e.what(); return -2; }
	return 0;
}
int ref_b200_download_state()
{
	TimeStepDFSPH_B200* ts = b200_ts();
	if (!ts) return -1;
	try { ts->downloadState(); } catch (const std::exception& e) { g_err = e.what(); return -2; }
	return 0;
}
int ref_b200_download_field(const char* name, void* dst)
{
	TimeStepDFSPH_B200* ts = b200_ts();
	if (!ts) return -1;
	try { ts->downloadField(name, static_cast<Real*>(dst)); } catch (const std::exception& e) { g_err = e.what(); return -2; }
	return 0;
}
/* Exercises the rest of the CompactNSearch surface of NeighborhoodSearch_B200 the way the reference uses it
   (add_point_set in model order, set_active, find_neighbors, z_sort + sort_field).  Returns 0 when every check holds,
   a positive code naming the first failed check otherwise. */
int ref_b200_facade_selftest()
{
	TimeStepDFSPH_B200* ts = b200_ts();
	if (!ts) return -1;
	Simulation* sim = Simulation::getCurrent();
	FluidModel* fm = sim->getFluidModel(0);
	const unsigned int n = fm->numActiveParticles();
	try {
		NeighborhoodSearch_B200 ns(*ts);
		if (ns.add_point_set(&fm->getPosition(0)[0], n, true, true, true, fm) != 0) return 1;          // FluidModel.cpp:323
		for (unsigned int b = 0; b < sim->numberOfBoundaryModels(); b++)
		{
			BoundaryModel_Akinci2012* bm = static_cast<BoundaryModel_Akinci2012*>(sim->getBoundaryModel(b));
			if (ns.add_point_set(&bm->getPosition(0)[0], bm->numberOfParticles(), false, false, true, bm) != b + 1) return 2;   // BoundaryModel_Akinci2012.cpp:109
		}
		if (ns.n_point_sets() != 1 + sim->numberOfBoundaryModels()) return 3;
		if (ns.point_set(0).get_user_data() != fm || ns.point_set(0).n_points() != n) return 4;
		if (!ns.is_active(0, 0) || (ns.n_point_sets() > 1 && (!ns.is_active(0, 1) || ns.is_active(1, 0) || ns.is_active(1, 1)))) return 5;
		ns.find_neighbors();
		unsigned long long total0 = 0, total1 = 0;
		for (unsigned int i = 0; i < n; i++)
		{
			const unsigned int c = ns.point_set(0).n_neighbors(0, i);
			const std::vector<unsigned int>& l = ns.point_set(0).neighbor_list(0, i);
			if (l.size() != c) return 6;
			for (unsigned int k = 0; k < c; k++) if (l[k] != ns.point_set(0).neighbor(0, i, k) || (k > 0 && l[k] <= l[k - 1])) return 7;
			total0 += c;
			if (ns.n_point_sets() > 1) total1 += ns.point_set(0).n_neighbors(1, i);
		}
		if (n > 1 && total0 == 0) return 8;
		// Simulation.cpp:713-754: switching a pair off empties its lists
		if (ns.n_point_sets() > 1)
		{
			ns.set_active(0u, 1u, false);
			for (unsigned int i = 0; i < n; i++) if (ns.point_set(0).n_neighbors(1, i) != 0) return 9;
			ns.set_active(0u, 1u, true);
			unsigned long long t = 0;
			for (unsigned int i = 0; i < n; i++) t += ns.point_set(0).n_neighbors(1, i);
			if (t != total1) return 10;
		}
		// z_sort + sort_field (Simulation.cpp:626, FluidModel.cpp:339-346): the permuted id field is the device order
		ns.z_sort();
		std::vector<unsigned int> ids(n);
		for (unsigned int i = 0; i < n; i++) ids[i] = i;
		ns.point_set(0).sort_field(ids.data());
		std::vector<char> seen(n, 0);
		for (unsigned int i = 0; i < n; i++) { if (ids[i] >= n || seen[ids[i]]) return 11; seen[ids[i]] = 1; }
		std::vector<Real> xs(n);
		for (unsigned int i = 0; i < n; i++) xs[i] = fm->getPosition(i)[0];
		ns.point_set(0).sort_field(xs.data());
		for (unsigned int i = 0; i < n; i++) if (xs[i] != fm->getPosition(ids[i])[0]) return 12;
	} catch (const std::exception& e) { g_err = e.what(); return -2; }
	return 0;
}

int ref_destroy()
{
	if (!Simulation::hasCurrent()) return 0;
	delete Simulation::getCurrent();   // also deletes TimeManager, models (+ their rigid bodies), time step (Simulation.cpp:91-110)
	g_step_seconds = 0.0;
	Utilities::Timing::m_averageTimes.clear();
	return 0;
}

}
