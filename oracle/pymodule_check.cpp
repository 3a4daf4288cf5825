// TEST INFRASTRUCTURE ONLY: a minimal stand-in for pySPlisHSPlasH's main.cpp, so that
// splishsplash_b200/host/DFSPH_B200Module.cpp can be compiled, linked against the reference stack (oracle/_ref) and
// imported on CPU (tests/test_dropin_cpu.py).  pySPlisHSPlasH itself cannot be built here (needs the full simulator).
#include <SPlisHSPlasH/TimeStep.h>
#include <pybind11/pybind11.h>

namespace py = pybind11;
void DFSPH_B200Module(py::module m_sub);

PYBIND11_MODULE(pydfsph_b200_check, m)
{
	// pySPlisHSPlasH registers TimeStep in TimeModule.cpp; the derived class needs its base to be known
	py::class_<SPH::TimeStep>(m, "TimeStep")
		.def("init", &SPH::TimeStep::init)
		.def("getMethodName", &SPH::TimeStep::getMethodName);
	DFSPH_B200Module(m);
}
