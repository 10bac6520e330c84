// Map-level stages of PostProcess.generate_poses that precede limb collection
// (reference decoder/factory.py:67-78, 98-146 and decoder/offset.py:8-43):
// flip-test fusion, x stride resize (bicubic A=-0.75 / bilinear, align_corners=False)
// and the optional scored_offset re-averaging.
//
// Rounding follows ATen's CPU kernels so that peak positions downstream are
// bit-identical to the reference run on CPU tensors: the generic interpolation loop
// accumulates  round(t1*w1) -> fma(t0, w0, .) -> fma(t2, w2, .) -> fma(t3, w3, .)
// along x and then along y (probed against torch 2.11, tests/golden/resize_small.npz).
// Interpolation weights are exact for power-of-two strides (the reference uses 4).
#include "og_common.cuh"
#include "og_interp.cuh"

namespace og {

namespace {

__global__ void resize_kernel(const float *__restrict__ in, float *__restrict__ out, int planes,
                              int h, int w, int scale, int cubic) {
    const int oh = h * scale, ow = w * scale;
    const float inv = 1.0f / (float)scale;
    const long long total = (long long)planes * oh * ow;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total;
         g += (long long)gridDim.x * blockDim.x) {
        const int ox = (int)(g % ow);
        const int oy = (int)((g / ow) % oh);
        const int pl = (int)(g / ((long long)ow * oh));
        const float *p = in + (size_t)pl * h * w;
        int ix[4], iy[4];
        float wx[4], wy[4];
        const int taps = axis_taps(ox, w, inv, cubic != 0, ix, wx);
        axis_taps(oy, h, inv, cubic != 0, iy, wy);
        float rows[4];
        for (int j = 0; j < taps; ++j) {
            float v[4];
            for (int i = 0; i < taps; ++i) v[i] = __ldg(p + iy[j] * w + ix[i]);
            rows[j] = combine(v, wx, taps);
        }
        out[g] = combine(rows, wy, taps);
    }
}

__global__ void flip_fuse_kernel(const float *__restrict__ hmp2n, const float *__restrict__ off2n,
                                 const FlipTablesDev ft, int n, int c, int l,
                                 int h, int w, float *__restrict__ out_hmp,
                                 float *__restrict__ out_off) {
    const long long hw = (long long)h * w;
    const long long n_h = (long long)n * c * hw;
    const long long n_o = (long long)n * 2 * l * hw;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < n_h + n_o;
         g += (long long)gridDim.x * blockDim.x) {
        if (g < n_h) {
            const int x = (int)(g % w);
            const int y = (int)((g / w) % h);
            const int ch = (int)((g / hw) % c);
            const int img = (int)(g / (hw * c));
            const float a = hmp2n[g];
            const float b = hmp2n[(((long long)(n + img) * c + ft.kp[ch]) * h + y) * w + (w - 1 - x)];
            out_hmp[g] = __fmul_rn(__fadd_rn(a, b), 0.5f);                  // factory.py:106
        } else {
            const long long q = g - n_h;
            const int x = (int)(q % w);
            const int y = (int)((q / w) % h);
            const int ch = (int)((q / hw) % (2 * l));
            const int img = (int)(q / (hw * 2 * l));
            const int limb = ch >> 1, comp = ch & 1;
            const float a = off2n[q];
            float r = a;                                                    // factory.py:134
            if (!((ft.reserved >> limb) & 1ull)) {
                float b = off2n[(((long long)(n + img) * 2 * l + 2 * (int)ft.limb[limb] + comp) * h + y) * w + (w - 1 - x)];
                if (comp == 0) b = -b;                                      // factory.py:132
                r = __fmul_rn(__fadd_rn(a, b), 0.5f);                       // factory.py:133
            }
            out_off[q] = r;
        }
    }
}

// out[n, 2l+comp] = sumpool_k(hmp[jf] * off) / (sumpool_k(hmp[jf]) + 1e-6)
__global__ void scored_offset_kernel(const float *__restrict__ hmp, const float *__restrict__ off,
                                     int n, int c, int l, int h, int w, int ksize, SkeletonDev sk,
                                     float *__restrict__ out) {
    const long long hw = (long long)h * w;
    const long long total = (long long)n * 2 * l * hw;
    const int pad = (ksize - 1) / 2;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total;
         g += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(g % w);
        const int y = (int)((g / w) % h);
        const int ch = (int)((g / hw) % (2 * l));
        const int img = (int)(g / (hw * 2 * l));
        const float *hp = hmp + ((long long)img * c + sk.from[ch >> 1]) * hw;
        const float *op = off + ((long long)img * 2 * l + ch) * hw;
        float s_w = 0.0f, s_wo = 0.0f;
        for (int yy = max(y - pad, 0); yy <= min(y + pad, h - 1); ++yy)
            for (int xx = max(x - pad, 0); xx <= min(x + pad, w - 1); ++xx) {
                const float hv = __ldg(hp + yy * w + xx);
                s_w = __fadd_rn(s_w, hv);
                s_wo = __fadd_rn(s_wo, __fmul_rn(hv, __ldg(op + yy * w + xx)));
            }
        out[g] = __fdiv_rn(s_wo, __fadd_rn(s_w, 1e-6f));
    }
}

__global__ void flip_average_kernel(const float *__restrict__ in2n, float *__restrict__ out, int n,
                                    int ch, int h, int w, ChannelPerm perm, int negate_even) {
    const long long hw = (long long)h * w;
    const long long total = (long long)n * ch * hw;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total;
         g += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(g % w);
        const int y = (int)((g / w) % h);
        const int c = (int)((g / hw) % ch);
        const int img = (int)(g / (hw * ch));
        float b = in2n[(((long long)(n + img) * ch + perm.src[c]) * h + y) * w + (w - 1 - x)];
        if (negate_even && (c & 1) == 0) b = -b;
        out[g] = __fmul_rn(__fadd_rn(in2n[g], b), 0.5f);
    }
}

__global__ void flip_cat_offsets_kernel(const float *__restrict__ off2n, float *__restrict__ out,
                                        int n, int l, int h, int w, ChannelPerm limb_flip,
                                        ChannelPerm reserved) {
    const long long hw = (long long)h * w;
    const long long total = (long long)n * 4 * l * hw;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total;
         g += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(g % w);
        const int y = (int)((g / w) % h);
        const int ch = (int)((g / hw) % (4 * l));
        const int img = (int)(g / (hw * 4 * l));
        const int limb = ch >> 2, comp = ch & 3;
        const float *orig = off2n + ((long long)img * 2 * l + 2 * limb + (comp & 1)) * hw;
        float r = orig[y * w + x];                       // comps 0, 1 and reserved limbs (:127)
        if (comp >= 2 && !reserved.src[limb]) {
            r = off2n[(((long long)(n + img) * 2 * l + 2 * limb_flip.src[limb] + (comp & 1)) * h + y) * w + (w - 1 - x)];
            if ((comp & 1) == 0) r = __fmul_rn(r, -1.0f);    // factory.py:124
        }
        out[g] = r;
    }
}

inline int grid_for(long long total, int threads) {
    long long b = (total + threads - 1) / threads;
    return (int)(b < 1 ? 1 : (b > 148LL * 64 ? 148LL * 64 : b));
}

}  // namespace

int launch_resize(const float *in, float *out, int planes, int h, int w, int scale, int mode,
                  cudaStream_t s) {
    const long long total = (long long)planes * h * scale * w * scale;
    if (total == 0) return OG_OK;
    resize_kernel<<<grid_for(total, 256), 256, 0, s>>>(in, out, planes, h, w, scale, mode);
    OG_CUDA_TRY(cudaGetLastError());
    return OG_OK;
}

int launch_flip_fuse(const float *hmp2n, const float *off2n, const FlipTablesDev &flips, int n, int c,
                     int l, int h, int w, float *out_hmp, float *out_off, cudaStream_t s) {
    const long long total = (long long)n * (c + 2 * l) * h * w;
    if (total == 0) return OG_OK;
    flip_fuse_kernel<<<grid_for(total, 256), 256, 0, s>>>(hmp2n, off2n, flips, n, c, l, h, w, out_hmp,
                                                         out_off);
    OG_CUDA_TRY(cudaGetLastError());
    return OG_OK;
}

int launch_flip_average(const float *in2n, float *out, int n, int ch, int h, int w,
                        const ChannelPerm &perm, bool negate_even, cudaStream_t s) {
    const long long total = (long long)n * ch * h * w;
    if (total == 0) return OG_OK;
    flip_average_kernel<<<grid_for(total, 256), 256, 0, s>>>(in2n, out, n, ch, h, w, perm,
                                                            negate_even ? 1 : 0);
    OG_CUDA_TRY(cudaGetLastError());
    return OG_OK;
}

int launch_flip_cat_offsets(const float *off2n, float *out, int n, int l, int h, int w,
                            const ChannelPerm &limb_flip, const ChannelPerm &reserved, cudaStream_t s) {
    const long long total = (long long)n * 4 * l * h * w;
    if (total == 0) return OG_OK;
    flip_cat_offsets_kernel<<<grid_for(total, 256), 256, 0, s>>>(off2n, out, n, l, h, w, limb_flip, reserved);
    OG_CUDA_TRY(cudaGetLastError());
    return OG_OK;
}

int launch_scored_offset(const float *hmp, const float *off, int n, int c, int l, int h, int w,
                         int ksize, const SkeletonDev &sk, float *out, cudaStream_t s) {
    const long long total = (long long)n * 2 * l * h * w;
    if (total == 0) return OG_OK;
    scored_offset_kernel<<<grid_for(total, 256), 256, 0, s>>>(hmp, off, n, c, l, h, w, ksize, sk, out);
    OG_CUDA_TRY(cudaGetLastError());
    return OG_OK;
}

}  // namespace og
