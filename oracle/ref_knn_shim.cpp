/*
 * ORACLE/_ref BUILD SHIM — TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Compiles the REFERENCE's own native KNN (nanoflann KD-tree) straight from
 * /root/reference/randlanet/utils/src/{neighbors.h,cloud.h,nanoflann.hpp},
 * included where they lie (-I on the command line, see oracle/Makefile), and
 * exposes it through a plain C ABI so that ctypes can call it without
 * libtorch.  This file contains no reference code: it only drives
 * nanoflann_knn_neighbors<float>() the way the reference's torch binding does:
 *   randlanet/utils/src/knn.cpp:43-61  per batch element, sequentially
 *   randlanet/utils/src/knn.cpp:15-17  error when support has fewer than k points
 *   randlanet/utils/src/knn.cpp:18-19  outputs pre-filled with -1
 *   randlanet/utils/src/knn.cpp:26-32  inputs copied into std::vector, then
 *                                      nanoflann_knn_neighbors (neighbors.h:281-322)
 * Output: int64 indices + SQUARED fp32 distances, ascending — same as knn_tpk.knn.
 * Single threaded, like the reference.
 */
#include <cstdint>
#include <vector>
#include "neighbors.h"

extern "C" int ref_knn_tpk(const float* support, const float* query, int B, int Ns,
                           int Nq, int K, int64_t* idx_out, float* d2_out)
{
    if (Ns < K) return -2;
    for (int b = 0; b < B; ++b) {
        std::vector<float> q(query + (size_t)b * Nq * 3, query + (size_t)(b + 1) * Nq * 3);
        std::vector<float> s(support + (size_t)b * Ns * 3, support + (size_t)(b + 1) * Ns * 3);
        std::vector<int64_t> nbr((size_t)Nq * K, -1);
        std::vector<float> dist((size_t)Nq * K, -1.0f);
        nanoflann_knn_neighbors<float>(q, s, nbr, dist, K);
        for (size_t i = 0; i < nbr.size(); ++i) {
            idx_out[(size_t)b * Nq * K + i] = nbr[i];
            d2_out[(size_t)b * Nq * K + i] = dist[i];
        }
    }
    return 0;
}
