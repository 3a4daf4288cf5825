// pybind11 exposure of the drop-in solver for pySPlisHSPlasH: the counterpart of
// pySPlisHSPlasH/DFSPHModule.cpp:50-66 for TimeStepDFSPH_B200 -- same static parameter handles, same methods -- plus the
// drop-in's own switches.  A maintainer adds this file to pySPlisHSPlasH/CMakeLists.txt and calls
// `DFSPH_B200Module(m_sub)` next to `DFSPHModule(m_sub)` in pySPlisHSPlasH/main.cpp.
// (SimulationDataDFSPH has no counterpart: the per-particle DFSPH fields live on the device and reach Python through
// the FluidModel field interface, "factor", "advected density", "p / rho^2", "p_v / rho^2", "pressure acceleration".)
#include <SPlisHSPlasH/TimeStep.h>
#include "TimeStepDFSPH_B200.h"

#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

namespace py = pybind11;

void DFSPH_B200Module(py::module m_sub)
{
	py::class_<SPH::TimeStepDFSPH_B200, SPH::TimeStep>(m_sub, "TimeStepDFSPH_B200")
		.def_readwrite_static("METHOD_NAME", &SPH::TimeStepDFSPH_B200::METHOD_NAME)
		.def_readwrite_static("SOLVER_ITERATIONS", &SPH::TimeStepDFSPH_B200::SOLVER_ITERATIONS)
		.def_readwrite_static("MIN_ITERATIONS", &SPH::TimeStepDFSPH_B200::MIN_ITERATIONS)
		.def_readwrite_static("MAX_ITERATIONS", &SPH::TimeStepDFSPH_B200::MAX_ITERATIONS)
		.def_readwrite_static("MAX_ERROR", &SPH::TimeStepDFSPH_B200::MAX_ERROR)
		.def_readwrite_static("SOLVER_ITERATIONS_V", &SPH::TimeStepDFSPH_B200::SOLVER_ITERATIONS_V)
		.def_readwrite_static("MAX_ITERATIONS_V", &SPH::TimeStepDFSPH_B200::MAX_ITERATIONS_V)
		.def_readwrite_static("MAX_ERROR_V", &SPH::TimeStepDFSPH_B200::MAX_ERROR_V)
		.def_readwrite_static("USE_DIVERGENCE_SOLVER", &SPH::TimeStepDFSPH_B200::USE_DIVERGENCE_SOLVER)

		.def(py::init<>())
		.def(py::init<const std::string&>(), py::arg("libraryPath"))
		.def("getMethodName", &SPH::TimeStepDFSPH_B200::getMethodName)
		.def("getNumIterations", &SPH::TimeStepDFSPH_B200::getNumIterations)
		.def("setSyncAllFields", &SPH::TimeStepDFSPH_B200::setSyncAllFields)
		.def("downloadNeighbors", [](SPH::TimeStepDFSPH_B200& ts, unsigned int other) {
			std::vector<unsigned int> offsets, indices;
			ts.downloadNeighbors(other, offsets, indices);
			return py::make_tuple(offsets, indices);   // CSR: row i = host particle i
		}, py::arg("other") = 0u);
}
