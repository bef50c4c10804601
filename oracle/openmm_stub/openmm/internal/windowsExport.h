#pragma once
#define OPENMM_EXPORT
