/* oracle/spmm_ref.c -- TEST INFRASTRUCTURE (see oracle/__init__.py).
 *
 * Plain-C restatement of the CPU SpMM that torch_sparse runs underneath
 * `matmul(adj_t, x, reduce=...)`, i.e. what SAGEConv / GCNConv execute at
 * /root/reference/plnlp/layer.py:20,23 when the reference runs on CPU.
 * torch_sparse is an un-vendored, un-pinned dependency of the reference
 * (README.md:15-19 pins only pyg 2.0.1); its published algorithm is:
 *   - parallel over output rows (upstream: at::parallel_for; here: pthreads pulling
 *     64-row blocks from an atomic counter -- libgomp is absent from this image),
 *   - per row, visit the stored entries in CSR order,
 *   - per output element, accumulate val*x (or x when value-less) in fp32,
 *     strictly in that order, one running sum per feature column,
 *   - reduce=mean divides the finished sum by max(row_nnz, 1).
 * Compiled with -ffp-contract=off so multiply and add are rounded separately,
 * as a baseline x86-64 build of the upstream extension does.
 *
 * rowptr: int64[M+1], col: int64[nnz] (upstream index width), val: float[nnz] or NULL.
 * x: [*, F] row-major with leading dimension ldx; out: [M, F] with ldo.
 * reduce: 0 = sum, 1 = mean.
 */
#include <stdint.h>
#include <stddef.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

typedef struct {
    const int64_t *rowptr, *col;
    const float *val, *x;
    void *out;
    int64_t ldx, ldo, M, F;
    int reduce, f64;
    atomic_llong next;
} job_t;

#define ROW_BLOCK 64

static void rows_f32(const job_t *j, int64_t m0, int64_t m1)
{
    const int64_t F = j->F;
    for (int64_t m = m0; m < m1; ++m) {
        const int64_t b = j->rowptr[m], e = j->rowptr[m + 1];
        float *o = (float *)j->out + m * j->ldo;
        for (int64_t k = 0; k < F; ++k) o[k] = 0.0f;
        for (int64_t p = b; p < e; ++p) {
            const float *xr = j->x + j->col[p] * j->ldx;
            if (j->val) {
                const float v = j->val[p];
                for (int64_t k = 0; k < F; ++k) o[k] += v * xr[k];
            } else {
                for (int64_t k = 0; k < F; ++k) o[k] += xr[k];
            }
        }
        if (j->reduce == 1) {
            const float cnt = (float)((e - b) > 0 ? (e - b) : 1);
            for (int64_t k = 0; k < F; ++k) o[k] = o[k] / cnt;
        }
    }
}

static void rows_f64(const job_t *j, int64_t m0, int64_t m1)
{
    const int64_t F = j->F;
    for (int64_t m = m0; m < m1; ++m) {
        const int64_t b = j->rowptr[m], e = j->rowptr[m + 1];
        double *o = (double *)j->out + m * j->ldo;
        for (int64_t k = 0; k < F; ++k) o[k] = 0.0;
        for (int64_t p = b; p < e; ++p) {
            const float *xr = j->x + j->col[p] * j->ldx;
            const double v = j->val ? (double)j->val[p] : 1.0;
            for (int64_t k = 0; k < F; ++k) o[k] += v * (double)xr[k];
        }
        if (j->reduce == 1) {
            const double cnt = (double)((e - b) > 0 ? (e - b) : 1);
            for (int64_t k = 0; k < F; ++k) o[k] = o[k] / cnt;
        }
    }
}

static void *worker(void *arg)
{
    job_t *j = (job_t *)arg;
    for (;;) {
        const int64_t m0 = atomic_fetch_add(&j->next, ROW_BLOCK);
        if (m0 >= j->M) break;
        const int64_t m1 = m0 + ROW_BLOCK < j->M ? m0 + ROW_BLOCK : j->M;
        if (j->f64) rows_f64(j, m0, m1); else rows_f32(j, m0, m1);
    }
    return NULL;
}

static void run(job_t *j, int threads)
{
    if (threads <= 0) threads = (int)sysconf(_SC_NPROCESSORS_ONLN);
    if (threads > 256) threads = 256;
    if (threads < 1) threads = 1;
    atomic_init(&j->next, 0);
    pthread_t tid[256];
    for (int t = 1; t < threads; ++t) pthread_create(&tid[t], NULL, worker, j);
    worker(j);
    for (int t = 1; t < threads; ++t) pthread_join(tid[t], NULL);
}

/* threads <= 0: use every online core. */
void plnlp_oracle_spmm_f32(const int64_t *rowptr, const int64_t *col, const float *val,
                           const float *x, int64_t ldx, float *out, int64_t ldo,
                           int64_t M, int64_t F, int reduce, int threads)
{
    job_t j = {rowptr, col, val, x, out, ldx, ldo, M, F, reduce, 0};
    run(&j, threads);
}

/* fp64-accumulating twin, used by the tests as a high-precision yardstick. */
void plnlp_oracle_spmm_f64(const int64_t *rowptr, const int64_t *col, const float *val,
                           const float *x, int64_t ldx, double *out, int64_t ldo,
                           int64_t M, int64_t F, int reduce, int threads)
{
    job_t j = {rowptr, col, val, x, out, ldx, ldo, M, F, reduce, 1};
    run(&j, threads);
}
