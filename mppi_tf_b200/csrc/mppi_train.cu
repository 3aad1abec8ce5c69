// Learner side of the MLP dynamics: one full-batch Adam step on the
// mean-squared error of the NORMALISED prediction, as LearnerBase._train_step does
// (/root/reference/scripts/src/learners/learner_base.py:469-496, optimizer tf.optimizers.Adam :325),
// for the network of mppi_mlp.cuh:   Xn = (concat(x, u) - Xmean)/Xstd,  Yn = ((x' - x) - Ymean)/Ystd,
//   h1 = relu(Xn W1 + b1), h2 = relu(h1 W2 + b2), out = h2 W3 + b3,  loss = mean((out - Yn)^2).
// fp32 accuracy throughout (the bf16 copy the rollout kernels stage is re-packed from the fp32 master weights after
// every step); the seven GEMMs of a step run on the tensor cores as error-compensated tf32 products (gemm_kernel).  This is
// the caller of the hot path, not the hot path: shared-memory-tiled kernels, fixed summation order (deterministic).
#include <cuda_bf16.h>
#include <mma.h>

#include "mppi_internal.h"
#include "mppi_mlp.cuh"

namespace mppi {

namespace {

constexpr int TS = 32;     // GEMM tile (M, N and K step)
constexpr int LDS_ = TS + 4;   // shared-memory leading dimension: a multiple of 4 floats (wmma), rows 16 apart stay 32-byte aligned

// C[M][N] = op(A) op(B) (+ bias[N]) (relu), row-major fp32.  TA: A is stored [K][M]; TB: B is stored [N][K].
// The contraction runs on the tensor cores: wmma m16n16k8 with tf32 operands and fp32 accumulation, each fp32 operand split
// exactly into three tf32 words (11 + 11 + 2 significant bits: the usual two-word "3xTF32" split keeps 22 of the 24 bits,
// four times the rounding of an fp32 product, and Adam turns the gradient error of nearly dead units into weight
// movement) and the six products above 2^-33 accumulated small to large - the learner works in fp32 because the
// reference's does (Keras in float64), and a plain tf32 or bf16 product would put 1e-3 into every gradient.  Four warps per CTA, one 16 x 16 fragment each; fixed
// summation order (deterministic).
template <bool TA, bool TB>
__global__ void __launch_bounds__(128) gemm_kernel(int M, int N, int K, const float *A, const float *B, const float *bias,
                                                   int relu, float *C)
{
    using namespace nvcuda;
    __shared__ __align__(32) float sA[TS][LDS_], sB[TS][LDS_], sC[TS][LDS_];
    const int tid = threadIdx.x, warp = tid >> 5;
    const int wm = warp >> 1, wn = warp & 1;                              // 2 x 2 warps
    const int m0 = blockIdx.y * TS, n0 = blockIdx.x * TS;
    // the tensor core's accumulator adds with truncation: a K-long chain drifts by about K 2^-24 relative.  Every 32-deep
    // tile is therefore accumulated from zero and added to the running sum with an ordinary (round-to-nearest) fp32 add.
    wmma::fragment<wmma::accumulator, 16, 16, 8, float> acc, part;
    wmma::fill_fragment(acc, 0.f);
    for (int k0 = 0; k0 < K; k0 += TS) {
        wmma::fill_fragment(part, 0.f);
        for (int i = tid; i < TS * TS; i += 128) {
            const int r = i / TS, c = i - r * TS;
            {   // sA[r][c] = op(A)[m0 + r][k0 + c]: consecutive threads walk the contiguous dimension of the source
                const int rr = TA ? c : r, cc = TA ? r : c;               // TA: source rows are k
                const int m = m0 + rr, k = k0 + cc;
                sA[rr][cc] = (m < M && k < K) ? (TA ? A[(size_t)k * M + m] : A[(size_t)m * K + k]) : 0.f;
            }
            {   // sB[r][c] = op(B)[k0 + r][n0 + c]
                const int rr = TB ? c : r, cc = TB ? r : c;               // TB: source rows are n
                const int k = k0 + rr, n = n0 + cc;
                sB[rr][cc] = (k < K && n < N) ? (TB ? B[(size_t)n * K + k] : B[(size_t)k * N + n]) : 0.f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < TS; kk += 8) {
            wmma::fragment<wmma::matrix_a, 16, 16, 8, wmma::precision::tf32, wmma::row_major> a_hi, a_lo, a_l2;
            wmma::fragment<wmma::matrix_b, 16, 16, 8, wmma::precision::tf32, wmma::row_major> b_hi, b_lo, b_l2;
            wmma::load_matrix_sync(a_hi, &sA[16 * wm][kk], LDS_);
            wmma::load_matrix_sync(b_hi, &sB[kk][16 * wn], LDS_);
#pragma unroll
            for (int i = 0; i < a_hi.num_elements; i++) {
                const float x = a_hi.x[i], hi = wmma::__float_to_tf32(x), lo = wmma::__float_to_tf32(x - hi);
                a_hi.x[i] = hi;
                a_lo.x[i] = lo;
                a_l2.x[i] = (x - hi) - lo;                                 // at most two significant bits: exact in tf32
            }
#pragma unroll
            for (int i = 0; i < b_hi.num_elements; i++) {
                const float x = b_hi.x[i], hi = wmma::__float_to_tf32(x), lo = wmma::__float_to_tf32(x - hi);
                b_hi.x[i] = hi;
                b_lo.x[i] = lo;
                b_l2.x[i] = (x - hi) - lo;
            }
            wmma::mma_sync(part, a_lo, b_lo, part);                        // 2^-22 terms
            wmma::mma_sync(part, a_l2, b_hi, part);
            wmma::mma_sync(part, a_hi, b_l2, part);
            wmma::mma_sync(part, a_lo, b_hi, part);                        // 2^-11 terms
            wmma::mma_sync(part, a_hi, b_lo, part);
            wmma::mma_sync(part, a_hi, b_hi, part);
        }
#pragma unroll
        for (int i = 0; i < acc.num_elements; i++) acc.x[i] += part.x[i];
        __syncthreads();
    }
    wmma::store_matrix_sync(&sC[16 * wm][16 * wn], acc, LDS_, wmma::mem_row_major);
    __syncthreads();
    for (int i = tid; i < TS * TS; i += 128) {
        const int r = i / TS, c = i - r * TS;
        const int m = m0 + r, n = n0 + c;
        if (m < M && n < N) {
            float v = sC[r][c] + (bias ? bias[n] : 0.f);
            if (relu) v = fmaxf(v, 0.f);
            C[(size_t)m * N + n] = v;
        }
    }
}

template <bool TA, bool TB>
cudaError_t gemm(int M, int N, int K, const float *A, const float *B, const float *bias, bool relu, float *C, cudaStream_t st)
{
    dim3 grid((N + TS - 1) / TS, (M + TS - 1) / TS);
    gemm_kernel<TA, TB><<<grid, 128, 0, st>>>(M, N, K, A, B, bias, relu ? 1 : 0, C);
    return cudaGetLastError();
}

// Xn[n][in] = (concat(x, u) - Xmean) / Xstd ; Yn[n][s] = ((x' - x) - Ymean) / Ystd
__global__ void prepare_kernel(int n, int s, int a, const float *x, const float *u, const float *xn, const float *norm,
                               float *Xn, float *Yn)
{
    const int in = s + a;
    const float *xmean = norm, *xinv = norm + 16, *ystd = norm + 32, *ymean = norm + 48;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n * in; i += gridDim.x * blockDim.x) {
        const int r = i / in, c = i - r * in;
        const float v = c < s ? x[(size_t)r * s + c] : u[(size_t)r * a + (c - s)];
        Xn[i] = (v - xmean[c]) * xinv[c];
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n * s; i += gridDim.x * blockDim.x) {
        const int c = i % s;
        Yn[i] = ((xn[i] - x[i]) - ymean[c]) / ystd[c];
    }
}

// dOut = 2 (out - Yn) / (n s); per-block partial sums of (out - Yn)^2 in a fixed order
__global__ void __launch_bounds__(256) loss_kernel(int total, const float *out, const float *Yn, float *dOut, float *partial)
{
    __shared__ float sred[256];
    float acc = 0.f;
    const float scale = 2.0f / (float)total;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const float d = out[i] - Yn[i];
        dOut[i] = scale * d;
        acc = fmaf(d, d, acc);
    }
    sred[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sred[threadIdx.x] += sred[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sred[0];
}
__global__ void loss_final_kernel(int nblocks, int total, const float *partial, float *loss)
{
    float acc = 0.f;
    for (int i = 0; i < nblocks; i++) acc += partial[i];
    *loss = acc / (float)total;
}

// d[i] = act[i] > 0 ? d[i] : 0   (ReLU backward)
__global__ void relu_mask_kernel(size_t total, const float *act, float *d)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
        if (!(act[i] > 0.f)) d[i] = 0.f;
}

// column sums g[c] = sum_r d[r][c], one thread per column, rows in order
__global__ void colsum_kernel(int n, int cols, const float *d, float *g)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    float acc = 0.f;
    for (int r = 0; r < n; r++) acc += d[(size_t)r * cols + c];
    g[c] = acc;
}

// Keras Adam (non-amsgrad): m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; w -= lr_t m / (sqrt(v) + eps),
// lr_t = lr sqrt(1 - b2^t) / (1 - b1^t)
__global__ void adam_kernel(int total, float *w, float *m, float *v, const float *g, float lr_t, float b1, float b2, float eps)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const float gi = g[i];
        const float mi = b1 * m[i] + (1.0f - b1) * gi;
        const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        w[i] -= lr_t * mi / (sqrtf(vi) + eps);
    }
}

// bf16 canonical blob from the fp32 master weights (device twin of mlp_pack_weights)
__global__ void pack_blob_kernel(int s, int a, const float *W1, const float *b1, const float *W2, const float *b2,
                                 const float *W3, const float *b3, uint8_t *blob)
{
    const int in = s + a;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < kWBlobBytes / 2; i += gridDim.x * blockDim.x)
        reinterpret_cast<__nv_bfloat16 *>(blob)[i] = __float2bfloat16(0.f);
    __threadfence();
    // a single CTA is launched: the zero fill above is complete for this CTA's own later writes after the barrier
    __syncthreads();
    auto put = [&](int off, int n, int k, int N, float v) {
        *reinterpret_cast<__nv_bfloat16 *>(blob + off + canon_offset_bytes(n, k, N)) = __float2bfloat16(v);
    };
    for (int i = threadIdx.x; i < kMlpH * (in + 1); i += blockDim.x) {
        const int n = i / (in + 1), k = i - n * (in + 1);
        put(0, n, k, kMlpH, k < in ? W1[k * kMlpH + n] : b1[n]);
    }
    for (int i = threadIdx.x; i < kMlpH * (kMlpH + 1); i += blockDim.x) {
        const int n = i / (kMlpH + 1), k = i - n * (kMlpH + 1);
        put(kW1Bytes, n, k, kMlpH, k < kMlpH ? W2[k * kMlpH + n] : b2[n]);
    }
    for (int i = threadIdx.x; i < s * (kMlpH + 1); i += blockDim.x) {
        const int n = i / (kMlpH + 1), k = i - n * (kMlpH + 1);
        put(kW1Bytes + kW2Bytes, n, k, kMlpNout, k < kMlpH ? W3[k * s + n] : b3[n]);
    }
}

}  // namespace

size_t mlp_param_count(int s, int a) { return (size_t)(s + a) * kMlpH + kMlpH + (size_t)kMlpH * kMlpH + kMlpH + (size_t)kMlpH * s + s; }

cudaError_t launch_pack_blob(int s, int a, const float *params, void *blob, cudaStream_t st)
{
    const int in = s + a;
    const float *W1 = params, *b1 = W1 + in * kMlpH, *W2 = b1 + kMlpH, *b2 = W2 + kMlpH * kMlpH, *W3 = b2 + kMlpH, *b3 = W3 + kMlpH * s;
    pack_blob_kernel<<<1, 1024, 0, st>>>(s, a, W1, b1, W2, b2, W3, b3, static_cast<uint8_t *>(blob));
    return cudaGetLastError();
}

// One Adam step.  params / m / v: the six tensors back to back (W1 b1 W2 b2 W3 b3, Keras layout);
// work: >= n * (in + 3 s + 4 H) + nparams + 1024 + 1 floats of scratch.  Returns the loss BEFORE the step
// in *loss_dev.
cudaError_t launch_train_step(int s, int a, int n, const float *x, const float *u, const float *xnext, const float *norm,
                              float *params, float *adam_m, float *adam_v, float lr_t, float b1c, float b2c, float eps,
                              float *work, float *loss_dev, cudaStream_t st)
{
    const int in = s + a, H = kMlpH;
    float *W1 = params, *b1 = W1 + in * H, *W2 = b1 + H, *b2 = W2 + H * H, *W3 = b2 + H, *b3 = W3 + H * s;
    const size_t np = mlp_param_count(s, a);
    float *Xn = work, *Yn = Xn + (size_t)n * in, *h1 = Yn + (size_t)n * s, *h2 = h1 + (size_t)n * H, *out = h2 + (size_t)n * H;
    float *dOut = out + (size_t)n * s, *dh2 = dOut + (size_t)n * s, *dh1 = dh2 + (size_t)n * H, *grad = dh1 + (size_t)n * H;
    float *partial = grad + np;
    float *gW1 = grad, *gb1 = gW1 + in * H, *gW2 = gb1 + H, *gb2 = gW2 + H * H, *gW3 = gb2 + H, *gb3 = gW3 + H * s;
    cudaError_t e;
    const int eb = 256, eg = 592;
    prepare_kernel<<<eg, eb, 0, st>>>(n, s, a, x, u, xnext, norm, Xn, Yn);
    // forward
    if ((e = gemm<false, false>(n, H, in, Xn, W1, b1, true, h1, st)) != cudaSuccess) return e;
    if ((e = gemm<false, false>(n, H, H, h1, W2, b2, true, h2, st)) != cudaSuccess) return e;
    if ((e = gemm<false, false>(n, s, H, h2, W3, b3, false, out, st)) != cudaSuccess) return e;
    const int nb = 1024;
    loss_kernel<<<nb, 256, 0, st>>>(n * s, out, Yn, dOut, partial);
    loss_final_kernel<<<1, 1, 0, st>>>(nb, n * s, partial, loss_dev);
    // backward
    if ((e = gemm<true, false>(H, s, n, h2, dOut, nullptr, false, gW3, st)) != cudaSuccess) return e;    // gW3 = h2^T dOut
    colsum_kernel<<<1, 32, 0, st>>>(n, s, dOut, gb3);
    if ((e = gemm<false, true>(n, H, s, dOut, W3, nullptr, false, dh2, st)) != cudaSuccess) return e;    // dh2 = dOut W3^T
    relu_mask_kernel<<<eg, eb, 0, st>>>((size_t)n * H, h2, dh2);
    if ((e = gemm<true, false>(H, H, n, h1, dh2, nullptr, false, gW2, st)) != cudaSuccess) return e;     // gW2 = h1^T dh2
    colsum_kernel<<<1, H, 0, st>>>(n, H, dh2, gb2);
    if ((e = gemm<false, true>(n, H, H, dh2, W2, nullptr, false, dh1, st)) != cudaSuccess) return e;     // dh1 = dh2 W2^T
    relu_mask_kernel<<<eg, eb, 0, st>>>((size_t)n * H, h1, dh1);
    if ((e = gemm<true, false>(in, H, n, Xn, dh1, nullptr, false, gW1, st)) != cudaSuccess) return e;    // gW1 = Xn^T dh1
    colsum_kernel<<<1, H, 0, st>>>(n, H, dh1, gb1);
    adam_kernel<<<64, 256, 0, st>>>((int)np, params, adam_m, adam_v, grad, lr_t, b1c, b2c, eps);
    return cudaGetLastError();
}

// augment_data of the learner (learner_base.py:455-467): every transition is repeated `samples` times and Gaussian noise of
// standard deviation sigma is added to the NORMALISED inputs; the normalised target is kept.  In raw coordinates:
// x' = x + d_x Xstd_x, u' = u + d_u Xstd_u, xnext' = xnext + (x' - x).  Noise: Philox4x32-10, counter (input column quad,
// augmented row, epoch, 0xA6), key = seed.  norm: the fvec block (Xmean [0,16), 1/Xstd [16,32), ...).
__global__ void augment_kernel(int n, int samples, int s, int a, const float *x, const float *u, const float *xnext, const float *norm,
                               float sigma, uint32_t key0, uint32_t key1, uint32_t epoch, float *xo, float *uo, float *xno)
{
    const int in = s + a, nq = (in + 3) >> 2;
    const long long total = (long long)n * samples * nq;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / nq;
        const int q = (int)(i - row * nq);
        const long long src = row / samples;
        float z[4];
        normals_from_words(philox4x32_10((uint32_t)q, (uint32_t)row, epoch, 0xA6u, key0, key1), z);
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int col = 4 * q + c;
            if (col >= in) break;
            const float d = sigma * z[c] / norm[16 + col];            // norm[16 + col] = 1 / Xstd
            if (col < s) {
                const float xv = x[src * s + col];
                xo[row * s + col] = xv + d;
                xno[row * s + col] = xnext[src * s + col] + ((xv + d) - xv);
            } else {
                uo[row * a + (col - s)] = u[src * a + (col - s)] + d;
            }
        }
    }
}

cudaError_t launch_augment(int n, int samples, int s, int a, const float *x, const float *u, const float *xnext, const float *norm,
                           float sigma, uint64_t seed, uint32_t epoch, float *xo, float *uo, float *xno, cudaStream_t st)
{
    augment_kernel<<<592, 256, 0, st>>>(n, samples, s, a, x, u, xnext, norm, sigma, (uint32_t)seed, (uint32_t)(seed >> 32), epoch, xo, uo, xno);
    return cudaGetLastError();
}

size_t train_work_floats(int s, int a, int n)
{
    return (size_t)n * ((s + a) + 3 * (size_t)s + 4 * (size_t)kMlpH) + mlp_param_count(s, a) + 1024 + 16;
}

}  // namespace mppi
