// GPU image preprocessing: the transform returned by the reference's `CLIPWrapper.get_preprocess()`
// (models/clip_wrapper.py:13,64-65 -> open_clip `image_transform`, inference mode) and applied per image by
// `dataset.py:31`:   Resize(R, BICUBIC) on the shorter side -> CenterCrop(R) -> ToTensor -> Normalize(mean, std).
//
// The arithmetic lives in a third-party dependency that is not part of the reference tree: Pillow's
// `src/libImaging/Resample.c` (ImagingResample, 8 bits per channel) behind torchvision's Resize.  Its published algorithm
// is restated here: per output coordinate a window [xmin, xmin+xmax) of the source, weights = bicubic (a = -0.5)
// evaluated at the tap centres with the support stretched by the down-scaling factor, normalised to sum 1 in double
// precision, rounded to 22-bit fixed point; a horizontal pass and a vertical pass, each accumulating in int32 from
// 1 << 21 and clipping to uint8 (so the intermediate image is rounded exactly like Pillow's).  Pinned against
// torchvision 0.26 / Pillow 12.2 outputs (tests/golden/preprocess_*.pt, tests/test_gpu_preprocess.py): bit-exact.
//
// Only the R x R crop window is computed: the horizontal pass produces the cropped columns of the source rows the
// vertical pass needs; the vertical pass writes normalised fp32 CHW directly.
#include "kernels.h"
#include "engine.h"
#include <cmath>
#include <map>
#include <tuple>
#include <vector>

namespace tapclip {
namespace {

constexpr int PRECISION_BITS = 32 - 8 - 2;

double bicubic_filter(double x) {
    const double a = -0.5;
    if (x < 0.0) x = -x;
    if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
    if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
    return 0.0;
}

struct Coeffs {
    int ksize = 0;
    std::vector<int> bounds;     // [out][2]: first source index, tap count
    std::vector<int> kk;         // [out][ksize] fixed-point weights
};

// Resample.c: precompute_coeffs + normalize_coeffs_8bpc for box = the whole axis
Coeffs precompute_coeffs(int in_size, int out_size) {
    Coeffs c;
    double scale = (double)in_size / out_size, filterscale = scale;
    if (filterscale < 1.0) filterscale = 1.0;
    const double support = 2.0 * filterscale;                       // bicubic support = 2
    c.ksize = (int)ceil(support) * 2 + 1;
    c.bounds.resize((size_t)out_size * 2);
    c.kk.assign((size_t)out_size * c.ksize, 0);
    std::vector<double> k(c.ksize);
    for (int xx = 0; xx < out_size; ++xx) {
        const double center = (xx + 0.5) * scale;
        double ww = 0.0;
        const double ss = 1.0 / filterscale;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        for (int x = 0; x < xmax; ++x) {
            const double w = bicubic_filter((x + xmin - center + 0.5) * ss);
            k[x] = w;
            ww += w;
        }
        for (int x = 0; x < xmax; ++x) {
            if (ww != 0.0) k[x] /= ww;
            c.kk[(size_t)xx * c.ksize + x] = k[x] < 0 ? (int)(-0.5 + k[x] * (1 << PRECISION_BITS)) : (int)(0.5 + k[x] * (1 << PRECISION_BITS));
        }
        c.bounds[(size_t)xx * 2] = xmin;
        c.bounds[(size_t)xx * 2 + 1] = xmax;
    }
    return c;
}

struct DevCoeffs { int ksize = 0; int* bounds = nullptr; int* kk = nullptr; std::vector<int> h_bounds; };
std::map<std::tuple<int, int, int>, DevCoeffs> g_coeffs;      // (device, in_size, out_size) -> device tables (one host thread per process by contract)

const DevCoeffs& device_coeffs(int in_size, int out_size, cudaStream_t st) {
    int dev;
    TC_CUDA(cudaGetDevice(&dev));
    auto key = std::make_tuple(dev, in_size, out_size);
    auto it = g_coeffs.find(key);
    if (it != g_coeffs.end()) return it->second;
    Coeffs c = precompute_coeffs(in_size, out_size);
    DevCoeffs d;
    d.ksize = c.ksize;
    d.h_bounds = c.bounds;
    TC_CUDA(cudaMalloc(&d.bounds, c.bounds.size() * sizeof(int)));
    TC_CUDA(cudaMalloc(&d.kk, c.kk.size() * sizeof(int)));
    TC_CUDA(cudaMemcpyAsync(d.bounds, c.bounds.data(), c.bounds.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    TC_CUDA(cudaMemcpyAsync(d.kk, c.kk.data(), c.kk.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    // first use of a size pair only: after this host-side wait the tables are complete for EVERY stream of the device (and the host
    // vectors may die with this scope)
    TC_CUDA(cudaStreamSynchronize(st));
    return g_coeffs.emplace(key, std::move(d)).first->second;
}

__device__ __forceinline__ uint8_t clip8(int v) {
    v >>= PRECISION_BITS;                                          // arithmetic shift, as Pillow's clip8 lookup index
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// tmp[r, c, ch] = horizontal resample of source row (row0 + r), output column (col0 + c);  src is [H, W, 3] uint8
__global__ void resample_h_kernel(const uint8_t* __restrict__ src, int W, uint8_t* __restrict__ tmp, int rows, int cols, int row0,
                                  int col0, const int* __restrict__ bounds, const int* __restrict__ kk, int ksize) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * cols) return;
    const int r = idx / cols, c = idx % cols;
    const int xx = col0 + c;
    const int xmin = bounds[xx * 2], xmax = bounds[xx * 2 + 1];
    const int* k = kk + (size_t)xx * ksize;
    const uint8_t* p = src + ((size_t)(row0 + r) * W + xmin) * 3;
    int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
    for (int x = 0; x < xmax; ++x) {
        const int w = k[x];
        s0 += p[x * 3] * w; s1 += p[x * 3 + 1] * w; s2 += p[x * 3 + 2] * w;
    }
    uint8_t* o = tmp + (size_t)idx * 3;
    o[0] = clip8(s0); o[1] = clip8(s1); o[2] = clip8(s2);
}

// out[ch, y, x] = ((vertical resample of tmp at output row top + y) / 255 - mean[ch]) / std[ch];  tmp is [rows, R, 3]
__global__ void resample_v_normalize_kernel(const uint8_t* __restrict__ tmp, int R, float* __restrict__ out, int top, int row0,
                                            const int* __restrict__ bounds, const int* __restrict__ kk, int ksize, float m0, float m1,
                                            float m2, float d0, float d1, float d2) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= R * R) return;
    const int y = idx / R, x = idx % R;
    const int yy = top + y;
    const int ymin = bounds[yy * 2] - row0, ymax = bounds[yy * 2 + 1];
    const int* k = kk + (size_t)yy * ksize;
    int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
    for (int t = 0; t < ymax; ++t) {
        const uint8_t* p = tmp + ((size_t)(ymin + t) * R + x) * 3;
        const int w = k[t];
        s0 += p[0] * w; s1 += p[1] * w; s2 += p[2] * w;
    }
    // torchvision ToTensor (uint8 -> float32 / 255) then Normalize ((x - mean) / std), all in fp32 with IEEE division
    out[idx] = __fdiv_rn(__fdiv_rn((float)clip8(s0), 255.f) - m0, d0);
    out[(size_t)R * R + idx] = __fdiv_rn(__fdiv_rn((float)clip8(s1), 255.f) - m1, d1);
    out[(size_t)2 * R * R + idx] = __fdiv_rn(__fdiv_rn((float)clip8(s2), 255.f) - m2, d2);
}

// horizontally resampled rows (uint8), grow-only, one scratch buffer per (device, stream): concurrent callers on different streams
// or devices never share it
std::map<std::pair<int, cudaStream_t>, DevBuf> g_tmp;

}  // namespace

void preprocess_image(const uint8_t* img, int H, int W, float* out, int R, int top, int left, const float* mean, const float* stdv,
                      cudaStream_t st) {
    TC_CHECK(H >= 1 && W >= 1 && R >= 1, "bad image size %dx%d -> %d", H, W, R);
    // torchvision _compute_resized_output_size for size = [R]: the shorter side becomes R, the longer int(R * long / short)
    int ow, oh;
    if (W <= H) { ow = R; oh = (int)((int64_t)R * H / W); } else { oh = R; ow = (int)((int64_t)R * W / H); }
    TC_CHECK(top >= 0 && left >= 0 && top + R <= oh && left + R <= ow, "crop window (%d,%d)+%d outside the resized image %dx%d", top, left, R, oh, ow);
    const DevCoeffs& ch = device_coeffs(W, ow, st);
    const DevCoeffs& cv = device_coeffs(H, oh, st);
    // source rows the cropped output rows touch
    const int row_first = cv.h_bounds[(size_t)top * 2];
    const int row_last = cv.h_bounds[(size_t)(top + R - 1) * 2] + cv.h_bounds[(size_t)(top + R - 1) * 2 + 1];
    const int rows = row_last - row_first;
    int dev;
    TC_CUDA(cudaGetDevice(&dev));
    DevBuf& tmp = g_tmp[{dev, st}];
    tmp.ensure((size_t)rows * R * 3);
    resample_h_kernel<<<(unsigned)ceil_div((int64_t)rows * R, 256), 256, 0, st>>>(img, W, (uint8_t*)tmp.p, rows, R, row_first, left,
                                                                             ch.bounds, ch.kk, ch.ksize);
    TC_LAUNCH_CHECK();
    resample_v_normalize_kernel<<<(unsigned)ceil_div((int64_t)R * R, 256), 256, 0, st>>>((const uint8_t*)tmp.p, R, out, top, row_first,
                                                                                    cv.bounds, cv.kk, cv.ksize, mean[0], mean[1], mean[2],
                                                                                    stdv[0], stdv[1], stdv[2]);
    TC_LAUNCH_CHECK();
}

}  // namespace tapclip
