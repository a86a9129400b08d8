// Node features and normalised targets of one sample in ONE launch -- data_utils.py:86-145,166-214:
//   wrap positions into the box, minimum-image frame differences -> velocities, normalise, flatten to x[N][3(W-1)+W],
//   and, when targets are given, y_acc[N][3] and y_temp_rate[N][1].
// Every operation is the single IEEE float32 operation the reference's CPU tensor ops perform, in the same order
// (explicit _rn intrinsics: no FMA contraction, true division), so the result is BIT-IDENTICAL to the reference's
// feature arithmetic -- also for inputs that already live on the GPU, where the same torch expressions would divide by
// a host scalar through a reciprocal multiply and differ in the last bit.
//   torch.remainder(a, b)      = fmod(a, b), plus b when the result is non-zero and its sign differs from b's
//   d[d < -B/2] += B; d[d > B/2] -= B   (strict comparisons against the float32 value of +-B/2)
#include "common.cuh"

namespace cgnn {
namespace {

constexpr int MAX_W = 16;

struct FeatMeta {
    float box, half_box, neg_half_box, dt;
    float vel_mean, vel_std, temp_mean, temp_std, acc_mean, acc_std, rate_mean, rate_std;
};

__device__ __forceinline__ float wrap_box(float a, float box) {
    float m = fmodf(a, box);
    if (m != 0.0f && ((box < 0.0f) != (m < 0.0f))) m = __fadd_rn(m, box);
    return m;
}
__device__ __forceinline__ float min_image(float d, const FeatMeta& md) {
    if (d < md.neg_half_box) d = __fadd_rn(d, md.box);
    if (d > md.half_box) d = __fsub_rn(d, md.box);
    return d;
}

// thread <-> particle; pos_seq[W][N][3] and temp_seq[W][N] are time-major (consecutive threads read consecutive particles)
__global__ void __launch_bounds__(256)
features_kernel(const float* __restrict__ pos_seq, const float* __restrict__ temp_seq, const float* __restrict__ pos_noise,
                const float* __restrict__ temp_noise, const float* __restrict__ target_pos, const float* __restrict__ target_temp,
                int64_t n, int W, FeatMeta md, float* __restrict__ recent_pos, float* __restrict__ x, float* __restrict__ y_acc,
                float* __restrict__ y_temp) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int F = 3 * (W - 1) + W;
    float* xr = x + i * F;
    float prev[3] = {0.f, 0.f, 0.f}, last_vel[3] = {0.f, 0.f, 0.f};
    float last_temp = 0.f;
    for (int t = 0; t < W; ++t) {
        float p[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float noise = pos_noise ? pos_noise[(i * W + t) * 3 + c] : 0.0f;
            p[c] = wrap_box(__fadd_rn(pos_seq[((int64_t)t * n + i) * 3 + c], noise), md.box);       // data_utils.py:92
        }
        if (t > 0) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float v = __fdiv_rn(min_image(__fsub_rn(p[c], prev[c]), md), md.dt);           // :100-107
                last_vel[c] = v;
                xr[(t - 1) * 3 + c] = __fdiv_rn(__fsub_rn(v, md.vel_mean), md.vel_std);              // :127-133
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) prev[c] = p[c];
        const float u = __fadd_rn(temp_seq[(int64_t)t * n + i], temp_noise ? temp_noise[i * W + t] : 0.0f);   // :95-97
        last_temp = u;
        xr[3 * (W - 1) + t] = __fdiv_rn(__fsub_rn(u, md.temp_mean), md.temp_std);                    // :135-141
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) recent_pos[i * 3 + c] = prev[c];
    if (target_pos != nullptr) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float noise = pos_noise ? pos_noise[(i * W + (W - 1)) * 3 + c] : 0.0f;
            const float tp = __fadd_rn(target_pos[i * 3 + c], noise);                                // :182
            const float nv = __fdiv_rn(min_image(__fsub_rn(tp, prev[c]), md), md.dt);                // :184-190
            const float a = __fdiv_rn(__fsub_rn(nv, last_vel[c]), md.dt);                            // :192
            y_acc[i * 3 + c] = __fdiv_rn(__fsub_rn(a, md.acc_mean), md.acc_std);                     // :195-197
        }
    }
    if (target_temp != nullptr) {
        const float tt = __fadd_rn(target_temp[i], temp_noise ? temp_noise[i * W + (W - 1)] : 0.0f); // :206
        const float r = __fdiv_rn(__fsub_rn(tt, last_temp), md.dt);                                  // :208
        y_temp[i] = __fdiv_rn(__fsub_rn(r, md.rate_mean), md.rate_std);                              // :212-214
    }
}

}  // namespace
}  // namespace cgnn

using namespace cgnn;

extern "C" int cgnn_preprocess_features(const float* pos_seq, const float* temp_seq, const float* pos_noise, const float* temp_noise,
                                        const float* target_pos, const float* target_temp, int64_t n, int32_t window, float box,
                                        float dt, const float* stats, float* recent_pos, float* x, float* y_acc, float* y_temp,
                                        cgnn_stream stream_) {
    CGNN_CHECK_ARG(pos_seq && temp_seq && stats && recent_pos && x && n >= 1, "cgnn_preprocess_features: bad arguments");
    CGNN_CHECK_ARG(window >= 2 && window <= MAX_W, "cgnn_preprocess_features: need 2 <= window <= %d (got %d)", MAX_W, window);
    CGNN_CHECK_ARG((target_pos == nullptr || y_acc != nullptr) && (target_temp == nullptr || y_temp != nullptr),
                   "cgnn_preprocess_features: a target needs its output");
    // +-box/2 as the reference forms them: Python doubles (-1 * box / 2, box / 2) rounded to float32 for the comparison
    FeatMeta md{box, (float)((double)box / 2.0), (float)(-1.0 * (double)box / 2.0), dt,
                stats[0], stats[1], stats[2], stats[3], stats[4], stats[5], stats[6], stats[7]};
    features_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(pos_seq, temp_seq, pos_noise, temp_noise, target_pos,
                                                                                   target_temp, n, window, md, recent_pos, x, y_acc, y_temp);
    CGNN_LAUNCH_CHECK();
    return CGNN_OK;
}
