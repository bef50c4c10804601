#pragma once
#include "openmm/KernelImpl.h"
