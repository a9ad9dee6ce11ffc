// TEST INFRASTRUCTURE — force-included (-include) when compiling the
// reference's sources, in place, for the CPU oracle.
//
// The reference hard-wires torch::kCUDA (src/solver.cpp:10, src/domain.cpp:7-11,
// src/colour.cpp:22-32, src/differential.hpp:19-46, src/ibm.cpp:13, every driver).
// This container and the GPU box's host side run the oracle on CPU, so after
// torch's own headers have been seen (they are include-guarded, so the
// reference's later #include <torch/torch.h> is a no-op) the identifier is
// re-pointed at kCPU.  No reference source is edited or copied.
#ifndef ORACLE_FORCE_CPU_HPP
#define ORACLE_FORCE_CPU_HPP
#include <torch/torch.h>
#define kCUDA kCPU
#endif
