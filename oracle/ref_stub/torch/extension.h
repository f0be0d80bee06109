/* Stand-in so that the reference's randlanet/utils/src/neighbors.h:7
 * (`#include <torch/extension.h>`; no torch symbol is used by that header)
 * resolves without libtorch when oracle/ref_knn_shim.cpp compiles the
 * reference sources where they lie.  The real header transitively provides
 * the std headers below, which neighbors.h relies on.  Test infrastructure only. */
#pragma once
#include <memory>
#include <vector>
#include <string>
#include <cstdint>
