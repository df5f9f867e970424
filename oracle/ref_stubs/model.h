// STUB for the un-vendored GCRANSAC "model.h" (danini/graph-cut-ransac): declares only the two names the reference's
// preemption_edge_length.h mentions in its signature and never touches.  Used only by oracle/ref_elc_wrap.cpp to
// compile that reference header UNMODIFIED, in place, as a checker for the oracle (oracle/Makefile target `ref`).
#pragma once
#include <cstddef>
#include <vector>
namespace gcransac
{
struct Model {};
struct Score {};
}  // namespace gcransac
