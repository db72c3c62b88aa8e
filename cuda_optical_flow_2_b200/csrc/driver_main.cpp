// driver_main.cpp -- headless host driver in the role of the reference's main.cu (main.cu:176-282),
// without camera, OpenCV or GUI: a synthetic translating frame sequence is pushed through
//   gpu::gauss_pyramid (main.cu:250) -> for k = levels-1..0: gpu::calc_opt_flow (main.cu:256-262)
//   -> flow composition (main.cu:136-147) -> prev/cur pyramid swap (main.cu:270-272)
// using the drop-in C++ wrappers of include/OptFlowGpuB200.hpp, and reports the median composed
// flow and the per-frame time.  `--batch N` runs the device-resident batched path instead
// (ofb_flow_pairs_host) on N pairs.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "OptFlowGpuB200.hpp"

static const float GAUS_KERNEL_3x3[9] = {0.0625f, 0.125f, 0.0625f, 0.125f, 0.25f, 0.125f, 0.0625f, 0.125f, 0.0625f};

// main.cu:95-104
template <typename T, int ch> static void alloc_pyramid(T ***pyramid, int w, int h, int levels)
{
    *pyramid = (T **)malloc(levels * sizeof(T *));
    for (int k = 0; k < levels; k++) {
        (*pyramid)[k] = (T *)calloc((size_t)w * h * ch, sizeof(T));
        w >>= 1;
        h >>= 1;
    }
}
template <typename T> static void free_pyramid(T ***pyramid, int levels)
{
    for (int i = 0; i < levels; i++) free((*pyramid)[i]);
    free(*pyramid);
}

// value-noise frame, 3 equal channels (the layout grayscale_avg leaves, OptFlowGpu.cu:58-59)
static void make_frame_c3(unsigned char *img, int w, int h, float dx, float dy, int cell, uint32_t seed)
{
    const int gw = w / cell + 4, gh = h / cell + 4;
    std::vector<float> g((size_t)gw * gh);
    uint32_t s = seed;
    for (auto &v : g) {
        s = s * 1664525u + 1013904223u;
        v = (float)((s >> 8) % 256u);
    }
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const float fx = ((float)x - dx) / cell + 1.5f, fy = ((float)y - dy) / cell + 1.5f;
            int ix = std::min(std::max((int)floorf(fx), 0), gw - 2), iy = std::min(std::max((int)floorf(fy), 0), gh - 2);
            const float ax = fx - ix, ay = fy - iy;
            float v = (1 - ay) * ((1 - ax) * g[iy * gw + ix] + ax * g[iy * gw + ix + 1]) +
                      ay * ((1 - ax) * g[(iy + 1) * gw + ix] + ax * g[(iy + 1) * gw + ix + 1]);
            const unsigned char c = (unsigned char)std::min(std::max(v, 0.f), 255.f);
            unsigned char *p = img + ((size_t)y * w + x) * 3;
            p[0] = p[1] = p[2] = c;
        }
}

static float median(std::vector<float> &v)
{
    if (v.empty()) return NAN;
    std::nth_element(v.begin(), v.begin() + v.size() / 2, v.end());
    return v[v.size() / 2];
}

int main(int argc, char **argv)
{
    int w = 640, h = 480, levels = 4, frames = 8, win = 19, warp = OFB_WARP_AS_WRITTEN, batch = 0;
    float step = 1.0f;
    const char *flo = nullptr; // --flo PREFIX: write PREFIX_<frame>.flo
    int arrow_res = 0;         // --arrows RES: count the arrows of visualizeFlowField(..., arrowRes)
    int solve = OFB_SOLVE_EXACT; // --solve 0|1: bit-exact or tolerance-mode 2x2 solve
    for (int i = 1; i < argc; i++) {
        auto arg = [&](const char *n) { return !strcmp(argv[i], n) && i + 1 < argc; };
        if (arg("--w")) w = atoi(argv[++i]);
        else if (arg("--h")) h = atoi(argv[++i]);
        else if (arg("--levels")) levels = atoi(argv[++i]);
        else if (arg("--frames")) frames = atoi(argv[++i]);
        else if (arg("--win")) win = atoi(argv[++i]);
        else if (arg("--warp")) warp = atoi(argv[++i]);
        else if (arg("--step")) step = (float)atof(argv[++i]);
        else if (arg("--batch")) batch = atoi(argv[++i]);
        else if (arg("--flo")) flo = argv[++i];
        else if (arg("--arrows")) arrow_res = atoi(argv[++i]);
        else if (arg("--solve")) solve = atoi(argv[++i]);
        else {
            fprintf(stderr, "usage: %s [--w W --h H --levels L --frames N --win WIN --warp 0|1|2 --solve 0|1 --step PX --batch N --flo PREFIX --arrows RES]\n", argv[0]);
            return 2;
        }
    }
    printf("Optical Flow (B200 path)\n=================\n%dx%d, %d levels, window %d, warp mode %d\n", w, h, levels, win, warp);
    gpu::set_lk_options(win, warp, 1.0f);
    if (!gpu::default_context()) return 1;
    gpu::set_lk_solve(solve);

    if (batch > 0) { // batched whole-pair path with host buffers
        ofb_params p{w, h, levels, win, warp, 1.0f, batch};
        std::vector<unsigned char> prev((size_t)w * h * 3 * batch), next(prev.size());
        for (int i = 0; i < batch; i++) {
            make_frame_c3(prev.data() + (size_t)i * w * h * 3, w, h, 0, 0, 8, 1234 + i);
            make_frame_c3(next.data() + (size_t)i * w * h * 3, w, h, step, 0.5f * step, 8, 1234 + i);
        }
        std::vector<std::vector<float>> flows(levels);
        std::vector<float *> fp(levels);
        for (int k = 0; k < levels; k++) {
            flows[k].resize((size_t)(w >> k) * (h >> k) * 2 * batch);
            fp[k] = flows[k].data();
        }
        for (int rep = 0; rep < 3; rep++) {
            auto t0 = std::chrono::steady_clock::now();
            int rc = ofb_flow_pairs_host(gpu::default_context(), &p, prev.data(), next.data(), 3, fp.data());
            auto t1 = std::chrono::steady_clock::now();
            if (rc) {
                fprintf(stderr, "ofb_flow_pairs_host: %s\n", ofb_last_error());
                return 1;
            }
            const double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
            printf("batch of %d pairs: %.2f ms  (%.1f Mpx-pairs/s end to end)\n", batch, ms, (double)w * h * batch / 1e6 / (ms / 1e3));
        }
        std::vector<float> us;
        for (size_t i = 0; i < flows[0].size(); i += 2)
            if (std::isfinite(flows[0][i])) us.push_back(flows[0][i]);
        printf("median residual u at level 0: %.4f\n", median(us));
        return 0;
    }

    unsigned char **prev_pyramid, **pyramid;
    float **flow_pyramid;
    alloc_pyramid<unsigned char, 3>(&prev_pyramid, w, h, levels);
    alloc_pyramid<unsigned char, 3>(&pyramid, w, h, levels);
    alloc_pyramid<float, 2>(&flow_pyramid, w, h, levels);

    std::vector<float> total((size_t)w * h * 2);
    make_frame_c3(prev_pyramid[0], w, h, 0, 0, 8, 1234);
    gpu::gauss_pyramid(prev_pyramid, w, h, levels, GAUS_KERNEL_3x3, 3, 3); // main.cu:209
    for (int f = 1; f <= frames; f++) {
        make_frame_c3(pyramid[0], w, h, step * f, 0.5f * step * f, 8, 1234);
        auto t0 = std::chrono::steady_clock::now();
        gpu::gauss_pyramid(pyramid, w, h, levels, GAUS_KERNEL_3x3, 3, 3); // main.cu:250
        for (int k = levels - 1; k >= 0; k--)                            // main.cu:256-262
            gpu::calc_opt_flow(prev_pyramid[k], pyramid[k], w >> k, h >> k, flow_pyramid, k, levels);
        auto t1 = std::chrono::steady_clock::now();
        if (gpu::last_status() != OFB_OK) return 1;
        // composition at level 0 (main.cu:136-147) on the device, in place of visualizeFlowField's per-arrow loop
        if (ofb_compose_flow_host(gpu::default_context(), flow_pyramid, w, h, levels, 0, total.data())) {
            fprintf(stderr, "ofb_compose_flow_host: %s\n", ofb_last_error());
            return 1;
        }
        std::vector<float> us, vs;
        for (size_t i = 0; i < total.size(); i += 2 * 49)
            if (std::isfinite(total[i]) && std::isfinite(total[i + 1])) {
                us.push_back(total[i]);
                vs.push_back(total[i + 1]);
            }
        printf("frame %d: %.2f ms, median composed flow (u,v) = (%.3f, %.3f) [units of 15/8 px, SURVEY Q1]\n", f,
               std::chrono::duration<double, std::milli>(t1 - t0).count(), median(us), median(vs));
        if (flo) { // export instead of imshow: the dense total flow as a Middlebury .flo file
            char path[1024];
            snprintf(path, sizeof path, "%s_%04d.flo", flo, f);
            if (ofb_write_flo(path, total.data(), w, h)) {
                fprintf(stderr, "ofb_write_flo: %s\n", ofb_last_error());
                return 1;
            }
        }
        if (arrow_res > 0) { // the arrows main.cu:125-171 would draw
            std::vector<int> arrows((size_t)4 * (arrow_res + 2) * (arrow_res + 2) * 4);
            int n = 0;
            if (ofb_flow_arrows_host(gpu::default_context(), flow_pyramid, w, h, levels, 0, arrow_res, arrows.data(),
                                     (int)(arrows.size() / 4), &n)) {
                fprintf(stderr, "ofb_flow_arrows_host: %s\n", ofb_last_error());
                return 1;
            }
            printf("  %d arrows on a grid of step %d px\n", n, w / arrow_res);
        }
        unsigned char **swap = prev_pyramid; // main.cu:270-272
        prev_pyramid = pyramid;
        pyramid = swap;
    }
    free_pyramid(&prev_pyramid, levels);
    free_pyramid(&pyramid, levels);
    free_pyramid(&flow_pyramid, levels);
    return 0;
}
