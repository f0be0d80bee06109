/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under 3d_recognizer_b200/ may
 * import, link or call this file; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it, as the checker.
 *
 * CPU restatement of the exact K-nearest-neighbour search that the reference's
 * only native component performs:
 *   randlanet/utils/src/knn.cpp:43-61      knn(): sequential loop over the batch
 *   randlanet/utils/src/knn.cpp:11-41      _single_batch_knn(): Ns >= k check, int64 idx + f32 d^2 out
 *   randlanet/utils/src/neighbors.h:281-322 nanoflann_knn_neighbors(): one knnSearch per query
 *   randlanet/utils/src/nanoflann.hpp:488-497  L2_Simple_Adaptor::evalMetric():
 *        result = 0; for dim in 0..2 { diff = a[dim]-b[dim]; result += diff*diff; }
 *        compiled per randlanet/CMakeLists.txt:23 with plain -O3 (x86-64 baseline,
 *        no FMA contraction) => d2 = fl(fl(fl(dx*dx) + fl(dy*dy)) + fl(dz*dz))
 *   randlanet/utils/src/nanoflann.hpp:152-233  KNNResultSet: K best kept ascending by d2
 *
 * What is restated and what is pinned:
 *   - the multiset of the K smallest d2 per query and their ascending order are
 *     what nanoflann returns; this file reproduces them bit-for-bit (checked
 *     against the compiled reference in oracle/_ref, tests/test_oracle.py).
 *   - nanoflann breaks exact d2 ties by KD-tree traversal order
 *     (nanoflann.hpp:1477-1491, no NANOFLANN_FIRST_MATCH).  The north-star
 *     contract is "ties broken by lower index", including at the K-th
 *     boundary, so the total order used here is (d2, index) ascending.  On
 *     tie-free rows the indices equal nanoflann's; on tied rows only d2 does.
 *
 * The scan order (ascending support index, strict '<' admission against the
 * current worst, stable insertion after equal keys) implements that total
 * order without ever comparing indices.
 *
 * Build: gcc -O2 -ffp-contract=off -pthread -shared -fPIC  (see oracle/Makefile)
 * -ffp-contract=off is REQUIRED: it pins the three separate roundings above.
 */
#include <stdint.h>
#include <stdlib.h>
#include <float.h>
#include <pthread.h>
#include <unistd.h>

#define ORACLE_KMAX 64
#define ORACLE_BLOCK 256

static inline float d2_exact(const float *q, const float *s)
{
    /* volatile-free but contraction is disabled by the build flags */
    const float dx = q[0] - s[0];
    const float dy = q[1] - s[1];
    const float dz = q[2] - s[2];
    float r = dx * dx;
    r = r + dy * dy;
    r = r + dz * dz;
    return r;
}

/* one query against one support cloud; out arrays have K entries */
static void knn_one_query(const float *support, int Ns, const float *q, int K,
                          int64_t *idx_out, float *d2_out)
{
    float bd[ORACLE_KMAX];
    int64_t bi[ORACLE_KMAX];
    int count = 0;
    float worst = FLT_MAX;
    float blk[ORACLE_BLOCK];

    for (int s0 = 0; s0 < Ns; s0 += ORACLE_BLOCK) {
        const int n = (Ns - s0 < ORACLE_BLOCK) ? (Ns - s0) : ORACLE_BLOCK;
        for (int j = 0; j < n; ++j)
            blk[j] = d2_exact(q, support + 3 * (size_t)(s0 + j));
        for (int j = 0; j < n; ++j) {
            const float d = blk[j];
            /* admission: while the list is not full everything enters; once it
             * is full only strictly smaller d2 enters, so an equal-d2 candidate
             * with a HIGHER index (we scan ascending) never evicts a lower one. */
            if (count == K && !(d < worst))
                continue;
            int pos = (count < K) ? count : K - 1;
            /* stable: move left only past strictly greater keys */
            while (pos > 0 && bd[pos - 1] > d) {
                bd[pos] = bd[pos - 1];
                bi[pos] = bi[pos - 1];
                --pos;
            }
            bd[pos] = d;
            bi[pos] = (int64_t)(s0 + j);
            if (count < K)
                ++count;
            if (count == K)
                worst = bd[K - 1];
        }
    }
    for (int k = 0; k < K; ++k) {
        idx_out[k] = bi[k];
        d2_out[k] = bd[k];
    }
}

struct job {
    const float *support, *query;
    int Ns, Nq, K;
    long long total;
    int64_t *idx_out;
    float *d2_out;
    long long next; /* work counter, advanced with an atomic add */
};

static void *worker(void *arg)
{
    struct job *j = (struct job *)arg;
    for (;;) {
        const long long t0 = __atomic_fetch_add(&j->next, 64, __ATOMIC_RELAXED);
        if (t0 >= j->total)
            break;
        const long long t1 = (t0 + 64 < j->total) ? t0 + 64 : j->total;
        for (long long t = t0; t < t1; ++t) {
            const int b = (int)(t / j->Nq);
            knn_one_query(j->support + (size_t)b * j->Ns * 3, j->Ns,
                          j->query + (size_t)t * 3, j->K,
                          j->idx_out + (size_t)t * j->K, j->d2_out + (size_t)t * j->K);
        }
    }
    return 0;
}

/*
 * support (B,Ns,3) f32, query (B,Nq,3) f32 row-major contiguous
 * idx_out (B,Nq,K) int64, d2_out (B,Nq,K) f32 (SQUARED distances, ascending)
 * returns 0, -1 bad argument, -2 Ns < K (knn.cpp:15-17 raises there)
 * threads <= 0 => all cores.  The reference itself is single threaded
 * (knn.cpp:53, no omp pragma); threads only speeds the checker up.
 */
int oracle_knn(const float *support, const float *query, int B, int Ns, int Nq,
               int K, int64_t *idx_out, float *d2_out, int threads)
{
    if (!support || !query || !idx_out || !d2_out)
        return -1;
    if (B < 0 || Ns < 0 || Nq < 0 || K <= 0 || K > ORACLE_KMAX)
        return -1;
    if (Ns < K)
        return -2;
    if (threads <= 0)
        threads = (int)sysconf(_SC_NPROCESSORS_ONLN);
    if (threads < 1)
        threads = 1;
    if (threads > 256)
        threads = 256;
    struct job j = {support, query, Ns, Nq, K, (long long)B * Nq, idx_out, d2_out, 0};
    if (threads == 1) {
        worker(&j);
        return 0;
    }
    pthread_t tid[256];
    int started = 0;
    for (int i = 0; i < threads - 1; ++i)
        if (pthread_create(&tid[started], 0, worker, &j) == 0)
            ++started;
    worker(&j);
    for (int i = 0; i < started; ++i)
        pthread_join(tid[i], 0);
    return 0;
}

int oracle_kmax(void) { return ORACLE_KMAX; }
