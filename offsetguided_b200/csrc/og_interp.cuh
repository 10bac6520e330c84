// Interpolation taps shared by the materialising resize kernel, the fused
// (resize + NMS) kernel and the fused offset sampling of K2.
//
// F.interpolate(scale_factor=S, align_corners=False) as ATen's generic CPU kernel
// evaluates it (reference decoder/factory.py:74-78): src = (dst + 0.5) / S - 0.5
// (clamped at 0 for bilinear), taps clamped to the image, cubic A = -0.75, and the
// accumulation order round(t1*w1) -> fma(t0, w0, .) -> fma(t2, w2, .) -> fma(t3, w3, .),
// probed bit-for-bit against torch 2.11 (tests/golden/resize_small.npz).
#pragma once

namespace og {

__device__ __forceinline__ void cubic_weights(float t, float w[4]) {
    const float A = -0.75f;
    const float x0 = t + 1.0f, x3 = 2.0f - t, x2 = 1.0f - t;
    w[0] = ((A * x0 - 5.0f * A) * x0 + 8.0f * A) * x0 - 4.0f * A;
    w[1] = ((A + 2.0f) * t - (A + 3.0f)) * t * t + 1.0f;
    w[2] = ((A + 2.0f) * x2 - (A + 3.0f)) * x2 * x2 + 1.0f;
    w[3] = ((A * x3 - 5.0f * A) * x3 + 8.0f * A) * x3 - 4.0f * A;
}

// Un-clamped first tap and weights of destination index `dst`; taps are
// first, first + 1, ... (2 for bilinear, 4 for bicubic).
__device__ __forceinline__ int axis_first_tap(int dst, float inv_scale, bool cubic, float w[4]) {
    float real = inv_scale * ((float)dst + 0.5f) - 0.5f;
    if (!cubic) real = fmaxf(real, 0.0f);
    const float fl = floorf(real);
    const float t = fminf(fmaxf(real - fl, 0.0f), 1.0f);
    if (cubic) {
        cubic_weights(t, w);
        return (int)fl - 1;
    }
    w[0] = 1.0f - t;
    w[1] = t;
    w[2] = w[3] = 0.0f;
    return (int)fl;
}

// taps clamped to [0, n_in); returns the tap count
__device__ __forceinline__ int axis_taps(int dst, int n_in, float inv_scale, bool cubic,
                                         int idx[4], float w[4]) {
    const int first = axis_first_tap(dst, inv_scale, cubic, w);
    const int taps = cubic ? 4 : 2;
#pragma unroll
    for (int j = 0; j < 4; ++j) idx[j] = min(max(first + j, 0), n_in - 1);
    return taps;
}

__device__ __forceinline__ float combine2(float v0, float v1, float w0, float w1) {
    return __fmaf_rn(v0, w0, __fmul_rn(v1, w1));
}
__device__ __forceinline__ float combine4(float v0, float v1, float v2, float v3, float w0, float w1,
                                          float w2, float w3) {
    float acc = __fmaf_rn(v0, w0, __fmul_rn(v1, w1));
    acc = __fmaf_rn(v2, w2, acc);
    return __fmaf_rn(v3, w3, acc);
}
__device__ __forceinline__ float combine(const float *v, const float *w, int taps) {
    float acc = __fmul_rn(v[1], w[1]);
    acc = __fmaf_rn(v[0], w[0], acc);
    for (int j = 2; j < taps; ++j) acc = __fmaf_rn(v[j], w[j], acc);
    return acc;
}

}  // namespace og
