#pragma once
namespace OpenMM {
inline int osrngseed() { return 12345; }   // the stand-in takes its noise from the driver, not from a seed
}
