// k_sessions.cu -- clip -> session aggregation, the step right after the extractor in the reference pipeline
// (src/utils.py:36-56 aggregate_clip_features: groupby('unique_participant_id').agg(['mean', 'std'])).
//
// One thread per (session, column) walks the session's clips in row order and reproduces the arithmetic of the pandas
// group kernels the reference ends up in, so results are bit-identical to the reference's:
//   mean : Kahan-compensated sum of the non-NaN values / their count            (pandas _libs.groupby.group_mean)
//   std  : Welford update  mean += (v - mean)/n;  M2 += (v - mean_new)*(v - mean_old);  sqrt(M2 / (n - 1)), NaN for n < 2
//                                                                               (pandas _libs.groupby.group_var, ddof = 1)
// The matrix is tiny (the paper's 866 clips x 25 columns -> 109 sessions x 50); the kernel exists so that the feature
// matrix never has to leave the device between extraction and aggregation.
#include "internal.h"
#include "common.cuh"

__global__ void k_session_agg(const double* __restrict__ feat, int n_cols, const int* __restrict__ row_start,
                              const int* __restrict__ rows, int n_groups, double* __restrict__ mean_out,
                              double* __restrict__ std_out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_groups * n_cols) return;
    const int g = t / n_cols, j = t % n_cols;
    double sumx = 0.0, comp = 0.0, mean = 0.0, m2 = 0.0;
    long long nobs = 0;
    for (int k = row_start[g]; k < row_start[g + 1]; k++) {
        const double v = feat[(size_t)rows[k] * n_cols + j];
        if (v == v) {
            nobs++;
            const double y = v - comp;
            const double s = sumx + y;
            comp = (s - sumx) - y;
            sumx = s;
            const double oldmean = mean;
            mean += (v - oldmean) / (double)nobs;
            m2 += (v - mean) * (v - oldmean);
        }
    }
    mean_out[(size_t)g * n_cols + j] = nobs > 0 ? sumx / (double)nobs : DEVNAN;
    std_out[(size_t)g * n_cols + j] = nobs > 1 ? sqrt(m2 / (double)(nobs - 1)) : DEVNAN;
}

void launch_session_agg(const double* feat, int n_cols, const int* row_start, const int* rows, int n_groups, double* mean_out,
                        double* std_out, cudaStream_t s) {
    const int total = n_groups * n_cols;
    if (total < 1) return;
    k_session_agg<<<(total + 127) / 128, 128, 0, s>>>(feat, n_cols, row_start, rows, n_groups, mean_out, std_out);
}
