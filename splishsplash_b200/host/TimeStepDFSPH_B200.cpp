#include "TimeStepDFSPH_B200.h"
#include "dfsph_b200.h"
#include "SPlisHSPlasH/TimeManager.h"
#include "SPlisHSPlasH/BoundaryModel_Akinci2012.h"
#include "SPlisHSPlasH/EmitterSystem.h"
#include "SPlisHSPlasH/AnimationFieldSystem.h"
#include "SPlisHSPlasH/Viscosity/Viscosity_Standard.h"
#include "Utilities/Timing.h"
#include "Utilities/Counting.h"
#include "Utilities/Logger.h"
#include <dlfcn.h>
#include <cstdlib>
#include <stdexcept>
#include <algorithm>
#include <cstring>

using namespace SPH;
using namespace GenParam;

std::string TimeStepDFSPH_B200::METHOD_NAME = "DFSPH_B200";
int TimeStepDFSPH_B200::SOLVER_ITERATIONS = -1;
int TimeStepDFSPH_B200::MIN_ITERATIONS = -1;
int TimeStepDFSPH_B200::MAX_ITERATIONS = -1;
int TimeStepDFSPH_B200::MAX_ERROR = -1;
int TimeStepDFSPH_B200::SOLVER_ITERATIONS_V = -1;
int TimeStepDFSPH_B200::MAX_ITERATIONS_V = -1;
int TimeStepDFSPH_B200::MAX_ERROR_V = -1;
int TimeStepDFSPH_B200::USE_DIVERGENCE_SOLVER = -1;

// function table resolved from the shared object (the binding a maintainer adds: see INTEGRATION.md)
struct TimeStepDFSPH_B200::Api
{
	decltype(&dfsph_b200_sizeof_real) sizeof_real;
	decltype(&dfsph_b200_default_config) default_config;
	decltype(&dfsph_b200_default_params) default_params;
	decltype(&dfsph_b200_create) create;
	decltype(&dfsph_b200_destroy) destroy;
	decltype(&dfsph_b200_last_error) last_error;
	decltype(&dfsph_b200_set_fluid) set_fluid;
	decltype(&dfsph_b200_add_boundary) add_boundary;
	decltype(&dfsph_b200_set_params) set_params;
	decltype(&dfsph_b200_step_host) step_host;
	decltype(&dfsph_b200_step) step;
	decltype(&dfsph_b200_download) download;
	decltype(&dfsph_b200_upload) upload;
	decltype(&dfsph_b200_neighbors) neighbors;
	decltype(&dfsph_b200_host_register) host_register;
	decltype(&dfsph_b200_host_unregister) host_unregister;
};

template <typename F>
static void resolve(void* lib, F& fn, const char* name)
{
	fn = reinterpret_cast<F>(dlsym(lib, name));
	if (!fn) throw std::runtime_error(std::string("TimeStepDFSPH_B200: symbol missing in CUDA library: ") + name);
}

void TimeStepDFSPH_B200::loadLibrary(const std::string& path)
{
	const std::string name = sizeof(Real) == 8 ? "libdfsph_b200_f64.so" : "libdfsph_b200_f32.so";
	std::string dir = path;
	if (dir.empty() && std::getenv("DFSPH_B200_LIB_DIR")) dir = std::getenv("DFSPH_B200_LIB_DIR");
	const std::string full = dir.empty() ? name : dir + "/" + name;
	m_lib = dlopen(full.c_str(), RTLD_NOW | RTLD_LOCAL);
	if (!m_lib) throw std::runtime_error("TimeStepDFSPH_B200: cannot load " + full + ": " + dlerror() + " (no CPU fallback)");
	m_api = new Api();
	resolve(m_lib, m_api->sizeof_real, "dfsph_b200_sizeof_real");
	resolve(m_lib, m_api->default_config, "dfsph_b200_default_config");
	resolve(m_lib, m_api->default_params, "dfsph_b200_default_params");
	resolve(m_lib, m_api->create, "dfsph_b200_create");
	resolve(m_lib, m_api->destroy, "dfsph_b200_destroy");
	resolve(m_lib, m_api->last_error, "dfsph_b200_last_error");
	resolve(m_lib, m_api->set_fluid, "dfsph_b200_set_fluid");
	resolve(m_lib, m_api->add_boundary, "dfsph_b200_add_boundary");
	resolve(m_lib, m_api->set_params, "dfsph_b200_set_params");
	resolve(m_lib, m_api->step_host, "dfsph_b200_step_host");
	resolve(m_lib, m_api->step, "dfsph_b200_step");
	resolve(m_lib, m_api->download, "dfsph_b200_download");
	resolve(m_lib, m_api->upload, "dfsph_b200_upload");
	resolve(m_lib, m_api->neighbors, "dfsph_b200_neighbors");
	resolve(m_lib, m_api->host_register, "dfsph_b200_host_register");
	resolve(m_lib, m_api->host_unregister, "dfsph_b200_host_unregister");
	if (m_api->sizeof_real() != (int)sizeof(Real)) throw std::runtime_error("TimeStepDFSPH_B200: Real size mismatch between the reference build and " + full);
}

void TimeStepDFSPH_B200::check(int rc, const char* what)
{
	if (rc == 0) return;
	const std::string msg = std::string("TimeStepDFSPH_B200: ") + what + " failed (" + std::to_string(rc) + "): " + m_api->last_error(m_ctx);
	LOG_ERR << msg;
	throw std::runtime_error(msg);
}

TimeStepDFSPH_B200::TimeStepDFSPH_B200(const std::string& libraryPath) :
	TimeStep(), m_lib(nullptr), m_ctx(nullptr), m_modelUploaded(false), m_api(nullptr)
{
	// defaults of TimeStepDFSPH (TimeStepDFSPH.cpp:33-41)
	m_iterations = 0;
	m_minIterations = 2;
	m_maxIterations = 100;
	m_maxError = static_cast<Real>(0.01);
	m_iterationsV = 0;
	m_enableDivergenceSolver = true;
	m_maxIterationsV = 100;
	m_maxErrorV = static_cast<Real>(0.1);
	m_syncAllFields = false;
	m_hostStateSync = true;
	m_hostStateStale = false;
	m_fieldDownloads = 0;
	m_uploadedParticles = 0;
	m_paramsPushed = false;
	for (int k = 0; k < NUM_MIRRORS; k++) m_stale[k] = false;
	static_assert(sizeof(dfsph_b200_params) <= sizeof(m_lastParams), "parameter cache too small");

	loadLibrary(libraryPath);
	resize();

	// the same particle fields as TimeStepDFSPH (TimeStepDFSPH.cpp:44-54), served from the host mirrors
	Simulation* sim = Simulation::getCurrent();
	if (sim->numberOfFluidModels() > 0)
	{
		FluidModel* model = sim->getFluidModel(0);
		model->addField({ "factor", METHOD_NAME, FieldType::Scalar, [this](const unsigned int i) -> Real* { return static_cast<Real*>(mirror(F_FACTOR, i)); } });
		model->addField({ "advected density", METHOD_NAME, FieldType::Scalar, [this](const unsigned int i) -> Real* { return static_cast<Real*>(mirror(F_DENSITY_ADV, i)); } });
		model->addField({ "p / rho^2", METHOD_NAME, FieldType::Scalar, [this](const unsigned int i) -> Real* { return static_cast<Real*>(mirror(F_KAPPA, i)); }, true });
		model->addField({ "p_v / rho^2", METHOD_NAME, FieldType::Scalar, [this](const unsigned int i) -> Real* { return static_cast<Real*>(mirror(F_KAPPA_V, i)); }, true });
		model->addField({ "pressure acceleration", METHOD_NAME, FieldType::Vector3, [this](const unsigned int i) -> Real* { return static_cast<Real*>(mirror(F_PRESSURE_ACCEL, i)); } });
	}
}

TimeStepDFSPH_B200::~TimeStepDFSPH_B200(void)
{
	Simulation* sim = Simulation::getCurrent();
	if (sim->numberOfFluidModels() > 0)
	{
		FluidModel* model = sim->getFluidModel(0);
		model->removeFieldByName("factor");
		model->removeFieldByName("advected density");
		model->removeFieldByName("p / rho^2");
		model->removeFieldByName("p_v / rho^2");
		model->removeFieldByName("pressure acceleration");
	}
	if (m_api) unpinHostArrays();
	if (m_ctx) m_api->destroy(m_ctx);
	delete m_api;
	if (m_lib) dlclose(m_lib);
}

void TimeStepDFSPH_B200::initParameters()
{
	TimeStep::initParameters();

	// identical names, groups and limits to TimeStepDFSPH::initParameters (TimeStepDFSPH.cpp:73-115)
	SOLVER_ITERATIONS = createNumericParameter("iterations", "Iterations", &m_iterations);
	setGroup(SOLVER_ITERATIONS, "Simulation|DFSPH");
	setDescription(SOLVER_ITERATIONS, "Iterations required by the pressure solver.");
	getParameter(SOLVER_ITERATIONS)->setReadOnly(true);

	MIN_ITERATIONS = createNumericParameter("minIterations", "Min. iterations", &m_minIterations);
	setGroup(MIN_ITERATIONS, "Simulation|DFSPH");
	setDescription(MIN_ITERATIONS, "Minimal number of iterations of the pressure solver.");
	static_cast<NumericParameter<unsigned int>*>(getParameter(MIN_ITERATIONS))->setMinValue(0);

	MAX_ITERATIONS = createNumericParameter("maxIterations", "Max. iterations", &m_maxIterations);
	setGroup(MAX_ITERATIONS, "Simulation|DFSPH");
	setDescription(MAX_ITERATIONS, "Maximal number of iterations of the pressure solver.");
	static_cast<NumericParameter<unsigned int>*>(getParameter(MAX_ITERATIONS))->setMinValue(1);

	MAX_ERROR = createNumericParameter("maxError", "Max. density error(%)", &m_maxError);
	setGroup(MAX_ERROR, "Simulation|DFSPH");
	setDescription(MAX_ERROR, "Maximal density error (%).");
	static_cast<RealParameter*>(getParameter(MAX_ERROR))->setMinValue(static_cast<Real>(1e-6));

	SOLVER_ITERATIONS_V = createNumericParameter("iterationsV", "Iterations (divergence)", &m_iterationsV);
	setGroup(SOLVER_ITERATIONS_V, "Simulation|DFSPH");
	setDescription(SOLVER_ITERATIONS_V, "Iterations required by the divergence solver.");
	getParameter(SOLVER_ITERATIONS_V)->setReadOnly(true);

	MAX_ITERATIONS_V = createNumericParameter("maxIterationsV", "Max. iterations (divergence)", &m_maxIterationsV);
	setGroup(MAX_ITERATIONS_V, "Simulation|DFSPH");
	setDescription(MAX_ITERATIONS_V, "Maximal number of iterations of the divergence solver.");
	static_cast<NumericParameter<unsigned int>*>(getParameter(MAX_ITERATIONS_V))->setMinValue(1);

	MAX_ERROR_V = createNumericParameter("maxErrorV", "Max. divergence error(%)", &m_maxErrorV);
	setGroup(MAX_ERROR_V, "Simulation|DFSPH");
	setDescription(MAX_ERROR_V, "Maximal divergence error (%).");
	static_cast<RealParameter*>(getParameter(MAX_ERROR_V))->setMinValue(static_cast<Real>(1e-6));

	USE_DIVERGENCE_SOLVER = createBoolParameter("enableDivergenceSolver", "Enable divergence solver", &m_enableDivergenceSolver);
	setGroup(USE_DIVERGENCE_SOLVER, "Simulation|DFSPH");
	setDescription(USE_DIVERGENCE_SOLVER, "Turn divergence solver on/off.");
}

void TimeStepDFSPH_B200::resize()
{
	// SimulationDataDFSPH::init (SimulationDataDFSPH.cpp:22-42): size the per-particle arrays from the fluid models
	Simulation* sim = Simulation::getCurrent();
	if (sim->numberOfFluidModels() > 1)
		throw std::runtime_error("TimeStepDFSPH_B200: multiphase scenes (more than one fluid model) are outside the B200 hot-path scope");
	if (sim->is2DSimulation())
		throw std::runtime_error("TimeStepDFSPH_B200: 2-D simulations are outside the B200 hot-path scope");
	const unsigned int n = sim->numberOfFluidModels() ? sim->getFluidModel(0)->numParticles() : 0;
	m_factor.assign(n, 0.0);
	m_density_adv.assign(n, 0.0);
	m_pressure_rho2.assign(n, 0.0);
	m_pressure_rho2_V.assign(n, 0.0);
	m_pressureAccel.assign(n, Vector3r::Zero());
	for (int k = 0; k < NUM_MIRRORS; k++) m_stale[k] = false;
	m_modelUploaded = false;
	if (m_api) unpinHostArrays();   // the model's arrays may be reallocated before the next upload
}

void TimeStepDFSPH_B200::reset()
{
	TimeStep::reset();
	resize();          // SimulationDataDFSPH::reset zeroes the warm-start values
	m_iterations = 0;
	m_iterationsV = 0;
}

void TimeStepDFSPH_B200::pushParameters()
{
	Simulation* sim = Simulation::getCurrent();
	dfsph_b200_params p;
	m_api->default_params(&p);
	p.time_step_size = TimeManager::getCurrent()->getTimeStepSize();
	const Real* g = sim->getVecValue<Real>(Simulation::GRAVITATION);
	for (int k = 0; k < 3; k++) p.gravitation[k] = g[k];
	p.min_iterations = m_minIterations;
	p.max_iterations = m_maxIterations;
	p.max_error = m_maxError;
	p.max_iterations_v = m_maxIterationsV;
	p.max_error_v = m_maxErrorV;
	p.enable_divergence_solver = m_enableDivergenceSolver ? 1 : 0;
	p.cfl_method = sim->getValue<int>(Simulation::CFL_METHOD);
	p.cfl_factor = sim->getValue<Real>(Simulation::CFL_FACTOR);
	p.cfl_min_time_step_size = sim->getValue<Real>(Simulation::CFL_MIN_TIMESTEPSIZE);
	p.cfl_max_time_step_size = sim->getValue<Real>(Simulation::CFL_MAX_TIMESTEPSIZE);
	// FluidModel "viscosityMethod": 0 = none, 1 = "Standard viscosity" runs on the device (next-row f1)
	FluidModel* fm = sim->getFluidModel(0);
	p.viscosity_method = (int)fm->getViscosityMethod();
	if (p.viscosity_method == 1)
	{
		NonPressureForceBase* v = fm->getViscosityBase();
		p.viscosity = v->getValue<Real>(Viscosity_Standard::VISCOSITY_COEFFICIENT);
		p.viscosity_boundary = v->getValue<Real>(Viscosity_Standard::VISCOSITY_COEFFICIENT_BOUNDARY);
	}
	// set_params synchronises the device and invalidates the solver-loop graphs: only call it when something changed
	if (m_paramsPushed && std::memcmp(&p, m_lastParams, sizeof(p)) == 0) return;
	check(m_api->set_params(m_ctx, &p), "dfsph_b200_set_params");
	std::memcpy(m_lastParams, &p, sizeof(p));
	m_paramsPushed = true;
}

void* TimeStepDFSPH_B200::mirror(int which, unsigned int i)
{
	if (m_stale[which] && m_ctx)
	{
		const unsigned int n = m_uploadedParticles;
		static const dfsph_b200_field ids[NUM_MIRRORS] = { DFSPH_B200_FIELD_FACTOR, DFSPH_B200_FIELD_DENSITY_ADV, DFSPH_B200_FIELD_KAPPA,
			DFSPH_B200_FIELD_KAPPA_V, DFSPH_B200_FIELD_PRESSURE_ACCEL };
		void* dst[NUM_MIRRORS] = { m_factor.data(), m_density_adv.data(), m_pressure_rho2.data(), m_pressure_rho2_V.data(),
			m_pressureAccel.empty() ? nullptr : &m_pressureAccel[0][0] };
		if (n > 0) check(m_api->download(m_ctx, ids[which], dst[which], (size_t)n * (which == F_PRESSURE_ACCEL ? 3 : 1) * sizeof(Real), 1), "download field mirror");
		m_stale[which] = false;
		m_fieldDownloads++;
	}
	switch (which)
	{
		case F_FACTOR: return &m_factor[i];
		case F_DENSITY_ADV: return &m_density_adv[i];
		case F_KAPPA: return &m_pressure_rho2[i];
		case F_KAPPA_V: return &m_pressure_rho2_V[i];
		default: return &m_pressureAccel[i][0];
	}
}

void TimeStepDFSPH_B200::downloadField(const std::string& name, Real* dst)
{
	Simulation* sim = Simulation::getCurrent();
	if (!m_ctx || sim->numberOfFluidModels() == 0) return;
	const unsigned int n = m_uploadedParticles;
	struct { const char* name; dfsph_b200_field id; int dim; } table[] = {
		{ "position", DFSPH_B200_FIELD_POSITION, 3 }, { "velocity", DFSPH_B200_FIELD_VELOCITY, 3 }, { "density", DFSPH_B200_FIELD_DENSITY, 1 },
		{ "factor", DFSPH_B200_FIELD_FACTOR, 1 }, { "advected density", DFSPH_B200_FIELD_DENSITY_ADV, 1 }, { "p / rho^2", DFSPH_B200_FIELD_KAPPA, 1 },
		{ "p_v / rho^2", DFSPH_B200_FIELD_KAPPA_V, 1 }, { "pressure acceleration", DFSPH_B200_FIELD_PRESSURE_ACCEL, 3 } };
	for (auto& t : table)
		if (name == t.name)
		{
			if (n > 0) check(m_api->download(m_ctx, t.id, dst, (size_t)n * t.dim * sizeof(Real), 1), "downloadField");
			m_fieldDownloads++;
			return;
		}
	throw std::runtime_error("TimeStepDFSPH_B200::downloadField: unknown field '" + name + "'");
}

void TimeStepDFSPH_B200::setHostStateSync(bool b)
{
	if (b && !m_hostStateSync) downloadState();
	m_hostStateSync = b;
}

void TimeStepDFSPH_B200::downloadState()
{
	if (!m_hostStateStale || !m_ctx) return;
	FluidModel* fm = Simulation::getCurrent()->getFluidModel(0);
	const unsigned int n = m_uploadedParticles;
	if (n > 0)
	{
		check(m_api->download(m_ctx, DFSPH_B200_FIELD_POSITION, &fm->getPosition(0)[0], (size_t)n * 3 * sizeof(Real), 1), "download position");
		check(m_api->download(m_ctx, DFSPH_B200_FIELD_VELOCITY, &fm->getVelocity(0)[0], (size_t)n * 3 * sizeof(Real), 1), "download velocity");
		check(m_api->download(m_ctx, DFSPH_B200_FIELD_DENSITY, &fm->getDensity(0), (size_t)n * sizeof(Real), 1), "download density");
	}
	m_hostStateStale = false;
}

void TimeStepDFSPH_B200::emittedParticles(FluidModel* model, const unsigned int startIndex)
{
	// the device copy no longer matches the model (particle reuse / new active particles): upload it again before the next step
	m_modelUploaded = false;
	unpinHostArrays();
}

void TimeStepDFSPH_B200::uploadModel()
{
	Simulation* sim = Simulation::getCurrent();
	if (sim->getBoundaryHandlingMethod() != BoundaryHandlingMethods::Akinci2012 && sim->numberOfBoundaryModels() > 0)
		throw std::runtime_error("TimeStepDFSPH_B200: only boundaryHandlingMethod 0 (Akinci2012) is supported on the B200 path");
	FluidModel* fm = sim->getFluidModel(0);
	if (fm->getViscosityMethod() > 1 || fm->getVorticityMethod() != 0 || fm->getDragMethod() != 0 ||
		fm->getSurfaceTensionMethod() != 0 || fm->getElasticityMethod() != 0)
		throw std::runtime_error("TimeStepDFSPH_B200: of the non-pressure forces only viscosityMethod 0 (none) and 1 (Standard viscosity) "
			"run on the B200 path; set vorticityMethod/dragMethod/surfaceTensionMethod/elasticityMethod to 0 (SURVEY.md H6)");

	// emitters and animation fields change particles on the host between steps (TimeStepDFSPH.cpp:240-241): not on this path
	if (fm->getEmitterSystem() != nullptr && fm->getEmitterSystem()->numEmitters() > 0)
		throw std::runtime_error("TimeStepDFSPH_B200: emitters are outside the B200 hot-path scope");
	if (sim->getAnimationFieldSystem() != nullptr && sim->getAnimationFieldSystem()->numAnimationFields() > 0)
		throw std::runtime_error("TimeStepDFSPH_B200: animation fields are outside the B200 hot-path scope");

	if (m_ctx) { m_api->destroy(m_ctx); m_ctx = nullptr; }
	m_paramsPushed = false;
	dfsph_b200_config cfg;
	m_api->default_config(&cfg);
	cfg.particle_radius = sim->getParticleRadius();
	cfg.kernel = sim->getKernel();          // "kernel" / "gradKernel" enum ids (Simulation.cpp:215-253)
	cfg.grad_kernel = sim->getGradKernel();
	if (const char* dev = std::getenv("DFSPH_B200_DEVICE")) cfg.device = std::atoi(dev);
	check(m_api->create(&cfg, &m_ctx), "dfsph_b200_create");
	pushParameters();

	const unsigned int n = fm->numActiveParticles();
	std::vector<unsigned int> state(n);
	for (unsigned int i = 0; i < n; i++) state[i] = static_cast<unsigned int>(fm->getParticleState(i));
	// host array index is the particle identity on the device side (row k of every by-id transfer = host slot k)
	check(m_api->set_fluid(m_ctx, n, n ? &fm->getPosition(0)[0] : nullptr, n ? &fm->getVelocity(0)[0] : nullptr, nullptr,
		state.data(), fm->getDensity0(), fm->getVolume(0)), "dfsph_b200_set_fluid");

	for (unsigned int b = 0; b < sim->numberOfBoundaryModels(); b++)
	{
		BoundaryModel_Akinci2012* bm = static_cast<BoundaryModel_Akinci2012*>(sim->getBoundaryModel(b));
		const bool dynamic = bm->getRigidBodyObject()->isDynamic() || bm->getRigidBodyObject()->isAnimated();
		const unsigned int nb = bm->numberOfParticles();
		check(m_api->add_boundary(m_ctx, nb, nb ? &bm->getPosition(0)[0] : nullptr, nb ? &bm->getVolume(0) : nullptr, dynamic ? 1 : 0),
			"dfsph_b200_add_boundary");
	}
	m_modelUploaded = true;
	m_uploadedParticles = n;
	m_hostStateStale = false;
	// page-lock the FluidModel's own x, v, density storage: step_host then copies at PCIe speed
	unpinHostArrays();
	if (n > 0)
	{
		void* arrays[3] = { &fm->getPosition(0)[0], &fm->getVelocity(0)[0], &fm->getDensity(0) };
		const size_t bytes[3] = { (size_t)fm->numParticles() * 3 * sizeof(Real), (size_t)fm->numParticles() * 3 * sizeof(Real), (size_t)fm->numParticles() * sizeof(Real) };
		for (int k = 0; k < 3; k++)
			if (m_api->host_register(arrays[k], bytes[k]) == 0) m_pinned.push_back(arrays[k]);
	}
}

void TimeStepDFSPH_B200::unpinHostArrays()
{
	for (void* p : m_pinned) m_api->host_unregister(p);
	m_pinned.clear();
}

void TimeStepDFSPH_B200::step()
{
	Simulation* sim = Simulation::getCurrent();
	TimeManager* tm = TimeManager::getCurrent();
	const Real h = tm->getTimeStepSize();
	if (sim->numberOfFluidModels() == 0) { tm->setTime(tm->getTime() + h); return; }
	FluidModel* fm = sim->getFluidModel(0);

	// a changed particle count (reset, state load, emitted particles) means the device copy is out of date
	if (m_modelUploaded && fm->numActiveParticles() != m_uploadedParticles) m_modelUploaded = false;
	if (!m_modelUploaded) uploadModel();
	pushParameters();

	// search + density + factor + divergence solve + CFL + pressure solve + advection, one call
	START_TIMING("DFSPH_B200 step");
	dfsph_b200_step_stats stats;
	const unsigned int n = fm->numActiveParticles();
	if (m_hostStateSync)
		check(m_api->step_host(m_ctx, n ? &fm->getPosition(0)[0] : nullptr, n ? &fm->getVelocity(0)[0] : nullptr,
			n ? &fm->getDensity(0) : nullptr, &stats), "dfsph_b200_step_host");
	else
	{
		check(m_api->step(m_ctx, &stats), "dfsph_b200_step");
		m_hostStateStale = true;
	}
	STOP_TIMING_AVG;

	m_iterations = stats.iterations;
	m_iterationsV = m_enableDivergenceSolver ? stats.iterations_v : 0;
	INCREASE_COUNTER("DFSPH - iterations", static_cast<Real>(m_iterations));      // TimeStepDFSPH.cpp:343
	if (m_enableDivergenceSolver)
		INCREASE_COUNTER("DFSPH - iterationsV", static_cast<Real>(m_iterationsV)); // :497

	// the five DFSPH fields: stale until somebody reads them (mirror()); eager mode fetches them now
	for (int k = 0; k < NUM_MIRRORS; k++) m_stale[k] = true;
	if (m_syncAllFields && n > 0)
		for (int k = 0; k < NUM_MIRRORS; k++) mirror(k, 0);

	// Simulation::updateTimeStepSize ran on the device (Simulation.cpp:395-493); publish its result, then advance time
	// with the step's initial h exactly like TimeStepDFSPH::step (:248)
	tm->setTimeStepSize(static_cast<Real>(stats.time_step_size));
	tm->setTime(tm->getTime() + h);
	// the device already holds this step size: the next pushParameters must not see it as a change
	reinterpret_cast<dfsph_b200_params*>(m_lastParams)->time_step_size = static_cast<Real>(stats.time_step_size);
}

void TimeStepDFSPH_B200::downloadNeighbors(unsigned int other, std::vector<unsigned int>& offsets, std::vector<unsigned int>& indices)
{
	Simulation* sim = Simulation::getCurrent();
	FluidModel* fm = sim->getFluidModel(0);
	const unsigned int n = fm->numActiveParticles();
	if (!m_modelUploaded) uploadModel();
	// host positions are authoritative: search exactly what the host sees
	if (n > 0 && !m_hostStateStale) check(m_api->upload(m_ctx, DFSPH_B200_FIELD_POSITION, &fm->getPosition(0)[0], n * 3 * sizeof(Real), 1), "upload position");
	std::vector<unsigned int> counts(n), ids(n);
	std::vector<uint64_t> off(n + 1, 0);
	check(m_api->neighbors(m_ctx, (int)other, counts.data(), nullptr, nullptr, 0), "dfsph_b200_neighbors");
	uint64_t total = 0;
	for (unsigned int i = 0; i < n; i++) total += counts[i];
	std::vector<unsigned int> idx(total ? total : 1);
	check(m_api->neighbors(m_ctx, (int)other, counts.data(), off.data(), idx.data(), total), "dfsph_b200_neighbors");
	if (n > 0) check(m_api->download(m_ctx, DFSPH_B200_FIELD_ID, ids.data(), n * sizeof(unsigned int), 0), "download id");
	// device rows -> host rows (id = host index); fluid neighbour indices -> host indices
	offsets.assign(n + 1, 0);
	for (unsigned int r = 0; r < n; r++) offsets[ids[r] + 1] = counts[r];
	for (unsigned int i = 0; i < n; i++) offsets[i + 1] += offsets[i];
	indices.resize(total);
	for (unsigned int r = 0; r < n; r++)
	{
		unsigned int* dst = indices.data() + offsets[ids[r]];
		for (unsigned int k = 0; k < counts[r]; k++) dst[k] = other == 0 ? ids[idx[off[r] + k]] : idx[off[r] + k];
		std::sort(dst, dst + counts[r]);
	}
}

const std::vector<unsigned int> NeighborhoodSearch_B200::s_empty;

unsigned int NeighborhoodSearch_B200::add_point_set(Real const* x, std::size_t n, bool is_dynamic, bool search_neighbors, bool find_neighbors, void* user_data)
{
	PointSet ps;
	ps.m_ns = this; ps.m_index = (unsigned int)m_sets.size(); ps.m_x = x; ps.m_n = n;
	ps.m_dynamic = is_dynamic; ps.m_search = search_neighbors; ps.m_find = find_neighbors; ps.m_user = user_data;
	m_sets.push_back(ps);
	for (auto& s : m_sets) s.m_ns = this;
	// CompactNSearch: a new set searches / is found according to its flags against every existing set
	const std::size_t k = m_sets.size();
	m_active.resize(k);
	for (std::size_t i = 0; i < k; i++) m_active[i].resize(k, false);
	for (std::size_t i = 0; i < k; i++)
	{
		m_active[i][k - 1] = m_sets[i].m_search && find_neighbors;
		m_active[k - 1][i] = search_neighbors && m_sets[i].m_find;
	}
	m_sortTable.resize(k);
	return ps.m_index;
}

void NeighborhoodSearch_B200::resize_point_set(unsigned int i, Real const* x, std::size_t n)
{
	m_sets[i].m_x = x; m_sets[i].m_n = n;
}

void NeighborhoodSearch_B200::set_active(bool active)
{
	for (auto& row : m_active) for (std::size_t j = 0; j < row.size(); j++) row[j] = active;
}

void NeighborhoodSearch_B200::set_active(unsigned int i, bool search_neighbors, bool find_neighbors)
{
	for (std::size_t j = 0; j < m_active.size(); j++) { m_active[i][j] = search_neighbors; m_active[j][i] = find_neighbors; }
}

void NeighborhoodSearch_B200::set_active(unsigned int i, unsigned int j, bool active)
{
	m_active[i][j] = active;
}

void NeighborhoodSearch_B200::find_neighbors()
{
	Simulation* sim = Simulation::getCurrent();
	const unsigned int nsets = 1 + sim->numberOfBoundaryModels();
	if (m_sets.empty())
	{
		// used without add_point_set (round-1 style): mirror the Simulation's sets with the reference's flags
		FluidModel* fm = sim->getFluidModel(0);
		add_point_set(&fm->getPosition(0)[0], fm->numActiveParticles(), true, true, true, fm);
		for (unsigned int b = 0; b < sim->numberOfBoundaryModels(); b++)
		{
			BoundaryModel_Akinci2012* bm = static_cast<BoundaryModel_Akinci2012*>(sim->getBoundaryModel(b));
			add_point_set(&bm->getPosition(0)[0], bm->numberOfParticles(), false, false, true, bm);
		}
	}
	std::vector<unsigned int> off, idx;
	m_off.assign(nsets, std::vector<unsigned int>());
	m_idx.assign(nsets, std::vector<unsigned int>());
	m_ts.downloadNeighbors(0, m_off[0], m_idx[0]);
	m_ts.downloadNeighbors(1, off, idx);
	// split the concatenated boundary lists into one CSR per boundary model (local indices)
	std::vector<unsigned int> bodyStart(1, 0u);
	for (unsigned int b = 0; b < sim->numberOfBoundaryModels(); b++)
		bodyStart.push_back(bodyStart.back() + static_cast<BoundaryModel_Akinci2012*>(sim->getBoundaryModel(b))->numberOfParticles());
	const std::size_t n = off.empty() ? 0 : off.size() - 1;
	for (unsigned int b = 0; b + 1 < nsets; b++)
	{
		std::vector<unsigned int>& o = m_off[b + 1]; std::vector<unsigned int>& x = m_idx[b + 1];
		o.assign(n + 1, 0u);
		for (std::size_t i = 0; i < n; i++)
		{
			for (unsigned int k = off[i]; k < off[i + 1]; k++)
				if (idx[k] >= bodyStart[b] && idx[k] < bodyStart[b + 1]) x.push_back(idx[k] - bodyStart[b]);
			o[i + 1] = (unsigned int)x.size();
		}
	}
}

void NeighborhoodSearch_B200::z_sort()
{
	m_ts.downloadSortTable(m_sortTable.empty() ? (m_sortTable.resize(1), m_sortTable[0]) : m_sortTable[0]);
}

unsigned int NeighborhoodSearch_B200::n_neighbors_of(unsigned int a, unsigned int set, unsigned int i) const
{
	if (a != 0 || set >= m_off.size() || !is_active(a, set) || m_off[set].empty()) return 0;
	return m_off[set][i + 1] - m_off[set][i];
}

unsigned int NeighborhoodSearch_B200::neighbor_of(unsigned int a, unsigned int set, unsigned int i, unsigned int k) const
{
	return m_idx[set][m_off[set][i] + k];
}

const std::vector<unsigned int>& NeighborhoodSearch_B200::list_of(unsigned int a, unsigned int set, unsigned int i) const
{
	if (n_neighbors_of(a, set, i) == 0) return s_empty;
	m_scratch.assign(m_idx[set].begin() + m_off[set][i], m_idx[set].begin() + m_off[set][i + 1]);
	return m_scratch;
}

void TimeStepDFSPH_B200::downloadSortTable(std::vector<unsigned int>& table)
{
	// device row r holds host particle id[r]: exactly the table PointSet::sort_field applies (new[r] = old[table[r]])
	const unsigned int n = m_uploadedParticles;
	table.assign(n, 0u);
	if (m_ctx && n > 0) check(m_api->download(m_ctx, DFSPH_B200_FIELD_ID, table.data(), (size_t)n * sizeof(unsigned int), 0), "download id");
}

void Simulation_B200::useB200Solver(const std::string& libraryPath)
{
	// what Simulation::setSimulationMethod does for every built-in method (Simulation.cpp:544-601)
	delete m_timeStep;
	m_timeStep = nullptr;
	m_simulationMethod = SimulationMethods::NumSimulationMethods;
	m_timeStep = new TimeStepDFSPH_B200(libraryPath);
	m_timeStep->init();
	setValue(Simulation::KERNEL_METHOD, Simulation::ENUM_KERNEL_PRECOMPUTED_CUBIC);
	setValue(Simulation::GRAD_KERNEL_METHOD, Simulation::ENUM_GRADKERNEL_PRECOMPUTED_CUBIC);
	if (m_simulationMethodChanged != nullptr)
		m_simulationMethodChanged();
}
