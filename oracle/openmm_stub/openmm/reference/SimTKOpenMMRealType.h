#pragma once
#include <cmath>
typedef double RealOpenMM;
#define EXP exp
#define SQRT sqrt
#define LOG log
#define POW pow
// OpenMM 7.x SimTKOpenMMRealType.h: Boltzmann constant in kJ/mol/K (same product the plugin's
// OpenCL side spells out at OpenCLSDMKernels.cpp:57-60)
#define BOLTZMANN (1.380658e-23)
#define AVOGADRO (6.0221367e23)
#define RGAS (BOLTZMANN * AVOGADRO)
#define BOLTZ (RGAS / 1000.0)
