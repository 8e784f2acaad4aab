// fft_aux.cuh - elementwise kernels either side of the power-of-two transform: Bluestein chirp
// multiply / pointwise product (algorithms/core/bluestein.c:107-148), r2c promotion / extraction
// (algorithms/auto/fft_auto.c:391-399), synthetic input fill. All are grid-stride, one 16-byte
// element per thread per iteration, fully coalesced; they are HBM-bound streaming kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fftb200 {

typedef double2 cd;

__device__ __forceinline__ cd cmul(cd a, cd b) {
    return make_double2(fma(a.x, b.x, -(a.y * b.y)), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ cd cmul_conj(cd a, cd b) {  // a * conj(b)
    return make_double2(fma(a.x, b.x, a.y * b.y), fma(a.y, b.x, -(a.x * b.y)));
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

// Counter-based synthetic input: element e -> (u(sm64(seed+2e)), u(sm64(seed+2e+1))),
// u(z) = (z >> 11) * 2^-52 - 1 in [-1, 1). Any slice can be regenerated on the host for parity checks.
__global__ void fill_splitmix_kernel(cd* dst, uint64_t seed, uint64_t first, uint64_t count) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        const uint64_t e = first + i;
        const double re = (double)(splitmix64(seed + 2 * e) >> 11) * (1.0 / 4503599627370496.0) - 1.0;
        const double im = (double)(splitmix64(seed + 2 * e + 1) >> 11) * (1.0 / 4503599627370496.0) - 1.0;
        dst[i] = make_double2(re, im);
    }
}

// y[i] = a[i] * b[i mod period]
__global__ void pointwise_mul_kernel(cd* y, const cd* a, const cd* b, size_t total, size_t period) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride)
        y[i] = cmul(a[i], __ldg(&b[i % period]));
}

// b[k] = chirp[k] for k < n, b[m-k] = chirp[k] for 1 <= k < n, else 0 (bluestein.c:116-121)
__global__ void bluestein_wrap_kernel(cd* b, const cd* chirp, int n, int m) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    cd v = make_double2(0.0, 0.0);
    if (i < n) v = chirp[i];
    else if (m - i < n) v = chirp[m - i];
    b[i] = v;
}

// work[t][k] = k < n ? x[t][k] * conj(chirp[k]) : 0 (bluestein.c:107-109)
__global__ void bluestein_pre_kernel(cd* work, const cd* x, const cd* chirp, int n, int m, size_t total) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const size_t t = i / (size_t)m;
        const int k = (int)(i - t * (size_t)m);
        cd v = make_double2(0.0, 0.0);
        if (k < n) v = cmul_conj(x[t * (size_t)n + k], __ldg(&chirp[k]));
        work[i] = v;
    }
}

// out[t][k] = work[t][k] * conj(chirp[k]) * scale (bluestein.c:139-148)
__global__ void bluestein_post_kernel(cd* out, const cd* work, const cd* chirp, int n, int m, size_t total,
                                      double scale) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const size_t t = i / (size_t)n;
        const int k = (int)(i - t * (size_t)n);
        cd v = cmul_conj(work[t * (size_t)m + k], __ldg(&chirp[k]));
        out[i] = make_double2(v.x * scale, v.y * scale);
    }
}

// real -> complex promotion (fft_auto.c:394-397)
__global__ void r2c_promote_kernel(cd* work, const double* x, size_t total) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride)
        work[i] = make_double2(x[i], 0.0);
}

// keep bins 0 .. n/2 of every transform (fft_auto.h:94: "Complex output array (size n/2+1)")
__global__ void r2c_extract_kernel(cd* out, const cd* work, size_t n, size_t nh, size_t total) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const size_t t = i / nh, k = i - t * nh;
        out[i] = work[t * n + k];
    }
}

// c2r, first step: rebuild the full Hermitian spectrum of a real signal from bins 0 .. n/2 (the inverse of
// r2c_extract: X[n - k] = conj(X[k])); the inverse c2c then runs on it with the reference's stage operators
__global__ void c2r_expand_kernel(cd* work, const cd* in, size_t n, size_t nh, size_t total) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const size_t t = i / n, k = i - t * n;
        cd v;
        if (k < nh) v = in[t * nh + k];
        else { v = in[t * nh + (n - k)]; v.y = -v.y; }
        work[i] = v;
    }
}

// c2r, last step: keep the real parts
__global__ void c2r_real_kernel(double* out, const cd* work, size_t total) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) out[i] = work[i].x;
}

// y[i] = conj(a[i]) * b[i] (cross-spectrum of applications/power_spectrum.c:176-178)
__global__ void pointwise_mul_conj_kernel(cd* y, const cd* a, const cd* b, size_t total) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) y[i] = cmul_conj(b[i], a[i]);
}

// dst[b][c][r] = src[b][r][c]: 32 x 32 tiles of 16-byte elements through shared memory (row of 33 elements: the
// column reads of the second phase fall into different banks); both phases move 512 contiguous bytes per warp.
// grid = (ceil(cols / 32), ceil(rows / 32), batch chunks), block = (32, 8).
__global__ void transpose_kernel(cd* __restrict__ dst, const cd* __restrict__ src, long long rows, long long cols, long long batch) {
    __shared__ cd tile[32][33];
    for (long long b = blockIdx.z; b < batch; b += gridDim.z) {
        const cd* s = src + b * rows * cols;
        cd* d = dst + b * rows * cols;
        for (long long r0 = (long long)blockIdx.y * 32; r0 < rows; r0 += (long long)gridDim.y * 32) {
            const long long c = (long long)blockIdx.x * 32 + threadIdx.x;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
                const long long r = r0 + threadIdx.y + j;
                if (r < rows && c < cols) tile[threadIdx.y + j][threadIdx.x] = s[r * cols + c];
            }
            __syncthreads();
            const long long r = r0 + threadIdx.x;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
                const long long cc = (long long)blockIdx.x * 32 + threadIdx.y + j;
                if (r < rows && cc < cols) d[cc * rows + r] = tile[threadIdx.x][threadIdx.y + j];
            }
            __syncthreads();
        }
    }
}

}  // namespace fftb200
