// hdsdp_b200/csrc/util.cu -- small HBM-bound helpers (copies, padding, mirroring).
#include "common.h"

namespace {

__global__ void set_identity_kernel(double *A, long lda, int n) {
    long idx = (long) blockIdx.x * blockDim.x + threadIdx.x;
    long total = (long) n * n;
    for (; idx < total; idx += (long) gridDim.x * blockDim.x) {
        int i = (int) (idx % n), j = (int) (idx / n);
        A[(long) j * lda + i] = (i == j) ? 1.0 : 0.0;
    }
}

// rows/cols in [n, np): identity on the diagonal, zero elsewhere (both the bottom rows and right columns)
__global__ void pad_identity_kernel(double *A, long lda, int n, int np) {
    long idx = (long) blockIdx.x * blockDim.x + threadIdx.x;
    int padw = np - n;
    long total = (long) np * padw * 2; // right columns (np x padw) + bottom rows (padw x np)
    for (; idx < total; idx += (long) gridDim.x * blockDim.x) {
        int i, j;
        if (idx < (long) np * padw) {
            i = (int) (idx % np);
            j = n + (int) (idx / np);
        } else {
            long t = idx - (long) np * padw;
            i = n + (int) (t % padw);
            j = (int) (t / padw);
        }
        A[(long) j * lda + i] = (i == j) ? 1.0 : 0.0;
    }
}

// A[j][i] (upper) <- A[i][j] (lower) with a 32x32 smem transpose; grid over lower tile pairs
__global__ void symmetrize_kernel(double *A, long lda, int n) {
    __shared__ double t[32][33];
    int bi = blockIdx.x, bj = blockIdx.y; // tile row, tile col ; only bi >= bj does work
    if (bi < bj) return;
    int tx = threadIdx.x, ty = threadIdx.y; // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        int i = bi * 32 + tx, j = bj * 32 + r;
        t[r][tx] = (i < n && j < n) ? A[(long) j * lda + i] : 0.0; // t[jj][ii] = A[i,j]
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        // write A[j, i] for j = bj*32 + tx (now the row), i = bi*32 + r (now the column)
        int j = bj * 32 + tx, i = bi * 32 + r;
        if (i < n && j < n && i > j) A[(long) i * lda + j] = t[tx][r];
    }
}

inline unsigned grid_for(long total, int threads) {
    long b = (total + threads - 1) / threads;
    long cap = (long) hd_num_sms() * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned) b;
}

__global__ void repack_kernel(double *__restrict__ dst, long ldd, const double *__restrict__ src, long lds, int n) {
    const long total = (long) n * n;
    for (long idx = (long) blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long) gridDim.x * blockDim.x) { // grid_for caps the grid
        const int i = (int) (idx % n), j = (int) (idx / n);
        dst[(long) j * ldd + i] = src[(long) j * lds + i];
    }
}
} // namespace

// Host <-> device transfer of an n x n column-major matrix whose device copy has leading dimension ld > n.  A 2-D copy between
// PAGEABLE host memory and the device degenerates into one small transfer per column (~13 us each: 2.7 ms for n = 200); instead
// the matrix travels contiguously through a device scratch buffer of n*n doubles and a kernel (un)packs it.
int hd_h2d_matrix(cudaStream_t st, double *d_dst, long ldd, const double *h_src, int n, double *d_stage) {
    if (ldd == n) { HD_CUDA(cudaMemcpyAsync(d_dst, h_src, sizeof(double) * (size_t) n * n, cudaMemcpyHostToDevice, st)); return HD_OK; }
    HD_CUDA(cudaMemcpyAsync(d_stage, h_src, sizeof(double) * (size_t) n * n, cudaMemcpyHostToDevice, st));
    HDK(repack_kernel)<<<grid_for((long) n * n, 256), 256, 0, st>>>(d_dst, ldd, d_stage, n, n);
    HD_CUDA(cudaGetLastError());
    return HD_OK;
}
int hd_d2h_matrix(cudaStream_t st, double *h_dst, const double *d_src, long lds, int n, double *d_stage) {
    if (lds == n) { HD_CUDA(cudaMemcpyAsync(h_dst, d_src, sizeof(double) * (size_t) n * n, cudaMemcpyDeviceToHost, st)); return HD_OK; }
    HDK(repack_kernel)<<<grid_for((long) n * n, 256), 256, 0, st>>>(d_stage, n, d_src, lds, n);
    HD_CUDA(cudaGetLastError());
    HD_CUDA(cudaMemcpyAsync(h_dst, d_stage, sizeof(double) * (size_t) n * n, cudaMemcpyDeviceToHost, st));
    return HD_OK;
}

int hd_set_identity(cudaStream_t st, double *A, long lda, int n) {
    HDK(set_identity_kernel)<<<grid_for((long) n * n, 256), 256, 0, st>>>(A, lda, n);
    HD_CUDA(cudaGetLastError());
    return HD_OK;
}

int hd_pad_identity(cudaStream_t st, double *A, long lda, int n, int np) {
    if (np == n) return HD_OK;
    HDK(pad_identity_kernel)<<<grid_for((long) np * (np - n) * 2, 256), 256, 0, st>>>(A, lda, n, np);
    HD_CUDA(cudaGetLastError());
    return HD_OK;
}

int hd_symmetrize_lower(cudaStream_t st, double *A, long lda, int n) {
    int t = (n + 31) / 32;
    HDK(symmetrize_kernel)<<<dim3(t, t), dim3(32, 8), 0, st>>>(A, lda, n);
    HD_CUDA(cudaGetLastError());
    return HD_OK;
}

int hd_copy2d(cudaStream_t st, double *dst, long ldd, const double *src, long lds, int rows, int cols) {
    if (rows <= 0 || cols <= 0) return HD_OK;
    if (ldd == rows && lds == rows) {
        HD_CUDA(cudaMemcpyAsync(dst, src, (size_t) rows * cols * sizeof(double), cudaMemcpyDeviceToDevice, st));
        return HD_OK;
    }
    HD_CUDA(cudaMemcpy2DAsync(dst, (size_t) ldd * 8, src, (size_t) lds * 8, (size_t) rows * 8, (size_t) cols,
                              cudaMemcpyDeviceToDevice, st));
    return HD_OK;
}
