// lk_level.cu -- host-side launcher of the fused per-level LK kernel: tensor maps, grid sizing,
// (window, warp mode) dispatch.
#include "lk_level.cuh"

#include <mutex>

namespace ofb {

PFN_encodeTiled get_encode_tiled()
{
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(sym);
    });
    return fn;
}

// u8 image batch as a 3-D tensor (x, y, image); box = LK_TILE_W x rows x 1; OOB reads give 0.
static int make_image_map(CUtensorMap *tm, const uint8_t *base, int w, int h, int n, size_t pitch, size_t stride,
                          int box_rows)
{
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return OFB_ERR_CUDA;
    }
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (pitch & 15) || (stride & 15)) {
        set_error("device images must be 16-byte aligned with pitch and image stride multiples of 16 "
                  "(base %p pitch %zu stride %zu)", (const void *)base, pitch, stride);
        return OFB_ERR_INVALID;
    }
    cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
    cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)(n > 1 ? stride : pitch * (size_t)h)};
    if (strides[1] & 15) strides[1] = (strides[1] + 15) & ~(cuuint64_t)15;
    cuuint32_t box[3] = {(cuuint32_t)LK_TILE_W, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (w %d h %d n %d pitch %zu)", (int)r, w, h, n, pitch);
        return OFB_ERR_CUDA;
    }
    return OFB_OK;
}

template <int WIN, int MODE>
static int launch_one(const LkLevelArgs &a, cudaStream_t stream, unsigned long long *launches)
{
    using C = LkCfg<WIN>;
    static bool attr_set[64] = {};
    int dev = 0;
    OFB_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 64 && !attr_set[dev]) {
        OFB_CUDA_TRY(cudaFuncSetAttribute(lk_level_kernel<WIN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          C::SMEM_BYTES));
        attr_set[dev] = true;
    }
    CUtensorMap tmP, tmQ;
    int rc = make_image_map(&tmP, a.prev, a.w, a.h_local, a.n_pairs, a.pitch, a.image_stride, C::CH);
    if (rc) return rc;
    rc = make_image_map(&tmQ, a.next, a.w, a.h_local, a.n_pairs, a.pitch, a.image_stride, C::CH);
    if (rc) return rc;

    const int out_rows = a.out_y1 - a.out_y0;
    const int strips = (a.w + C::TWO - 1) / C::TWO;
    // Rows per CTA.  The hardware hands CTAs to SMs as slots free up, so a launch takes about
    // (total row-steps) / SMs plus a ragged tail of roughly half a CTA's lifetime (MIN_BLOCKS CTAs
    // share an SM, so a CTA lives MIN_BLOCKS times its own row-steps).  Short CTAs shrink the tail,
    // tall CTAs amortise the 2R+2 halo rows: take the split that minimises the sum.
    const int n_sm = a.sm_count > 0 ? a.sm_count : 148;
    const long long cols = (long long)strips * a.n_pairs;
    const int max_ny = (out_rows + C::CH - 1) / C::CH;
    int rows_per_block = out_rows;
    double best = 1e300;
    for (int ny = 1; ny <= max_ny && ny <= 128; ny++) {
        int rpb = (out_rows + ny - 1) / ny;
        rpb = ((rpb + C::SUB - 1) / C::SUB) * C::SUB;
        const int nb = (out_rows + rpb - 1) / rpb;
        const double steps = rpb + 2 * C::R + 2 + C::SUB;
        const double cost = (double)(cols * nb) * steps / n_sm + 0.5 * C::MIN_BLOCKS * steps;
        if (cost < best * 0.999) {
            best = cost;
            rows_per_block = rpb;
        }
    }
    const int nby = (out_rows + rows_per_block - 1) / rows_per_block;

    LkKernelParams p;
    p.next = a.next;
    p.pitch = (int)a.pitch;
    p.image_stride = a.image_stride;
    p.w = a.w;
    p.h_local = a.h_local;
    p.y_off = a.y_off;
    p.h_global = a.h_global;
    p.out_y0 = a.out_y0;
    p.out_y1 = a.out_y1;
    p.rows_per_block = rows_per_block;
    p.as_written = (a.warp_mode == OFB_WARP_AS_WRITTEN) ? 1 : 0;
    p.scale2 = 2.0f * a.flow_scale;
    p.scale512 = 512.0f * a.flow_scale;
    p.cum_in = reinterpret_cast<const float2 *>(a.cum_in);
    p.cum_w = a.cum_w;
    p.cum_h_global = a.cum_h_global;
    p.cum_y_off = a.cum_y_off;
    p.cum_h_local = a.cum_h_local;
    p.cum_pair_stride = a.cum_pair_stride;
    p.flow_out = reinterpret_cast<float2 *>(a.flow_out);
    p.cum_out = reinterpret_cast<float2 *>(a.cum_out);
    p.flow_pair_stride = a.flow_pair_stride;
    p.reach_overflow = a.reach_overflow;

    dim3 grid((unsigned)strips, (unsigned)nby, (unsigned)a.n_pairs);
    lk_level_kernel<WIN, MODE><<<grid, LK_NT, C::SMEM_BYTES, stream>>>(tmP, tmQ, p);
    OFB_CUDA_TRY(cudaGetLastError());
    if (launches) ++*launches;
    return OFB_OK;
}

template <int WIN> static int launch_mode(const LkLevelArgs &a, cudaStream_t s, unsigned long long *l)
{
    if (a.cum_in == nullptr) return launch_one<WIN, 0>(a, s, l);
    if (a.warp_mode == OFB_WARP_BILINEAR) return launch_one<WIN, 2>(a, s, l);
    return launch_one<WIN, 1>(a, s, l);
}

int launch_lk_level(const LkLevelArgs &a, cudaStream_t stream, unsigned long long *launches)
{
    if (a.w < 1 || a.h_local < 1 || a.n_pairs < 1 || a.out_y0 < 0 || a.out_y1 > a.h_local || a.out_y0 >= a.out_y1) {
        set_error("lk_level: bad geometry (w %d h_local %d pairs %d out rows [%d,%d))", a.w, a.h_local, a.n_pairs,
                  a.out_y0, a.out_y1);
        return OFB_ERR_INVALID;
    }
    if (a.pitch * (size_t)a.h_local > 0x7fffffffull || (size_t)a.w * a.h_local > 0x3fffffffull) {
        set_error("lk_level: one image is limited to 2^31 bytes (pitch %zu x %d rows)", a.pitch, a.h_local);
        return OFB_ERR_INVALID;
    }
    if (a.n_pairs > 65535) {
        set_error("lk_level: at most 65535 pairs per launch (got %d)", a.n_pairs);
        return OFB_ERR_INVALID;
    }
    if (a.warp_mode < OFB_WARP_AS_WRITTEN || a.warp_mode > OFB_WARP_BILINEAR) {
        set_error("lk_level: unknown warp mode %d", a.warp_mode);
        return OFB_ERR_INVALID;
    }
    if ((reinterpret_cast<uintptr_t>(a.flow_out) & 15) || (a.cum_out && (reinterpret_cast<uintptr_t>(a.cum_out) & 15))) {
        set_error("lk_level: flow buffers must be 16-byte aligned");
        return OFB_ERR_INVALID;
    }
    switch (a.win) {
    case 3: return launch_mode<3>(a, stream, launches);
    case 5: return launch_mode<5>(a, stream, launches);
    case 7: return launch_mode<7>(a, stream, launches);
    case 9: return launch_mode<9>(a, stream, launches);
    case 11: return launch_mode<11>(a, stream, launches);
    case 13: return launch_mode<13>(a, stream, launches);
    case 15: return launch_mode<15>(a, stream, launches);
    case 17: return launch_mode<17>(a, stream, launches);
    case 19: return launch_mode<19>(a, stream, launches);
    default:
        set_error("lk_level: window %d not supported (odd 3..19)", a.win);
        return OFB_ERR_UNSUPPORTED;
    }
}

} // namespace ofb
