// raster_stream.cuh -- round-2 raster pipeline: coarse binning, pair-parallel forward that SAVES one
// compact record per surviving (pixel, face) pair, and a streaming backward that consumes them.
// Included by raster.cu (same translation unit: -fmad=false, same exact-twin arithmetic helpers).
//
// Why (profiles/r01_raster_v6_bwd_pairs_ncu_summary.txt): the round-1 backward
// re-derived fragment() (~13 IEEE divisions + a double sigmoid) for every pair the forward had already
// evaluated, behind 3 CTA barriers per 32-face chunk, at 0.9 % of the HBM roofline with DRAM 99 % idle;
// and every tile CTA re-scanned all F cull boxes in both passes.  Now:
//   k_bin_coarse     one CTA per 64x64-pixel bin scans the image's F cull boxes ONCE (TMA-staged, ordered
//                    ballot compaction) -> ascending face list per bin.  Tiles scan their bin's list
//                    (~100-200 entries) instead of F (1280 / 5120).
//   k_raster_fwd2    one CTA per 16x16 tile.  Per sub-chunk of faces:
//                      phase A (pair-parallel, all lanes busy): every candidate (pixel, face) of the faces'
//                        cull-box rectangles runs the exact-twin fragment(); survivors leave D / depth /
//                        texel id in a shared slot AND, when a pair buffer is given, one 48-byte record in HBM
//                        (32-candidate blocks, survivors compacted to the block front, block-SoA so a warp
//                        writes three fully coalesced 512-byte lines);
//                      phase B (pixel-parallel): each pixel walks the faces in ascending index (ordered
//                        semantics: p2f prefix-max weights kernel.cu:421-430, hard z-buffer strict '<' :409)
//                        and folds its slots into alpha / softmax / colour; p2f by warp-shuffle reduction.
//   k_raster_bwd2    one CTA per tile streams the tile's blocks: no cull boxes, no fragment(), no barrier
//                    after the per-pixel inputs are staged; a warp owns a contiguous block range, so the 9
//                    vertex gradients are accumulated privately and flushed (shuffle reduce + 9 RED) once per
//                    (warp, face run).
// Tiles whose blocks do not fit the caller's pair buffer are marked UNSAVED and take the round-1 recompute
// kernel (k_raster_bwd_pairs) -- results are identical either way (tests/test_raster_stream_gpu.py).
#pragma once

namespace umr {

constexpr int CB = 64;             // coarse bin side in pixels (4 x 4 tiles)
constexpr int LCAP = 512;          // coarse-list window == longest tile-list segment held in shared memory
#ifndef UMR_SUB_BLOCKS
#define UMR_SUB_BLOCKS 64
#endif
#ifndef UMR_FWD2_CTAS
#define UMR_FWD2_CTAS 3
#endif
constexpr int SUB_BLOCKS = UMR_SUB_BLOCKS;     // 32-candidate blocks per sub-chunk
constexpr int SLOTS = SUB_BLOCKS * 32;
constexpr uint32_t SEG_NONE = 0xffffffffu;
constexpr int32_t TILE_EMPTY = -1, TILE_UNSAVED = -2;
constexpr int BLK_F4 = 96;         // float4 per block: 3 planes x 32 records

// slot / record flag bits
constexpr uint32_t SL_VALID = 1u << 31, SL_ZV = 1u << 30, SL_FRONT = 1u << 29, SL_INS = 1u << 28, SL_TIX = 0xffffu;

// c_rcpw[w] = ceil(65536 / w): (l * c_rcpw[w]) >> 16 == l / w exactly for l < 1024, w <= 32
__constant__ uint32_t c_rcpw[33] = {0,     65536, 32768, 21846, 16384, 13108, 10923, 9363, 8192, 7282, 6554,
                                    5958,  5462,  5042,  4682,  4370,  4096,  3856,  3641, 3450, 3277, 3121,
                                    2979,  2850,  2731,  2622,  2521,  2428,  2341,  2260, 2185, 2115, 2048};

struct PairBuf {
    uint32_t* ctrl;      // [0] block cursor (== blocks wanted, may exceed cap), [1] tiles left unsaved
    int32_t* tile_head;  // [B * tiles]: first segment (block index), TILE_EMPTY or TILE_UNSAVED
    int32_t* ulist;      // [B * tiles]: ids of the unsaved tiles (count = ctrl[1]), walked by the recompute fallback
    uint32_t* blk_hdr;   // [cap]: per block  face | count << 16 ; per segment  [base] = #blocks, [base+1] = next
    float4* recs;        // [cap][3][32]
    uint32_t cap;        // blocks
};

struct PairBufLayout {
    size_t ctrl_off, head_off, ulist_off, hdr_off, rec_off, total;
};
inline PairBufLayout pair_layout(int B, int S, size_t cap_blocks) {
    PairBufLayout L;
    const size_t nt = (size_t)((S + TILE - 1) / TILE) * ((S + TILE - 1) / TILE) * B;
    L.ctrl_off = 0;
    L.head_off = 256;
    L.ulist_off = L.head_off + align256(nt * sizeof(int32_t));
    L.hdr_off = L.ulist_off + align256(nt * sizeof(int32_t));
    L.rec_off = L.hdr_off + align256(cap_blocks * sizeof(uint32_t));
    L.total = L.rec_off + cap_blocks * (size_t)BLK_F4 * sizeof(float4);
    return L;
}
// largest capacity (blocks) that fits `bytes`
inline size_t pair_capacity(int B, int S, size_t bytes) {
    const PairBufLayout z = pair_layout(B, S, 0);
    if (bytes <= z.total + 512) return 0;
    size_t cap = (bytes - z.total - 512) / ((size_t)BLK_F4 * sizeof(float4) + sizeof(uint32_t));
    while (cap > 0 && pair_layout(B, S, cap).total > bytes) --cap;
    return cap;
}

// ---------------------------------------------------------------------------------------------
// coarse binning: grid (ncb, ncb, B).  clist[(b, cy, cx)][F] u16, ccount[(b, cy, cx)]
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CTA) k_bin_coarse(const float4* __restrict__ box_all, const uint32_t* __restrict__ ubox,
                                                    uint16_t* __restrict__ clist, int* __restrict__ ccount, int F, int S) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4* s_box = reinterpret_cast<float4*>(smem_raw);
    __shared__ uint64_t s_bar;
    __shared__ int s_warp_cnt[NWARP];
    __shared__ float s_ext[4];
    const int b = blockIdx.z;
    if (threadIdx.x == 0) {
        mbar_init(&s_bar, 1);
        fence_mbar_init();
    }
    tile_extents(S, s_ext, CB);
    __syncthreads();
    const size_t cidx = ((size_t)b * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    int n = 0;
    uint32_t bar_phase = 0;
    if (!tile_outside_union(ubox, b, s_ext))
        n = build_tile_list(box_all + (size_t)b * F, F, s_ext[0], s_ext[1], s_ext[2], s_ext[3], s_box, clist + cidx * F,
                            s_warp_cnt, &s_bar, bar_phase);
    if (threadIdx.x == 0) ccount[cidx] = n;
}

// ---------------------------------------------------------------------------------------------
// visible faces: which faces win at least one pixel of the hard z-buffer (kernel.cu:404-415) -- the [B,F] bytes TexCycle
// derives from the hard render's face-index plane (loss_utils.py:161-166).  FACE-parallel: one CTA per 64x64 bin keeps a
// z-buffer of packed (depth bits << 32 | face) keys in shared memory; each warp takes faces of the bin's list and tests only
// the pixels of the face's bounding box (+2 px; the whole cull box for thin faces, R_FLG bit 4), atomicMin keeps the
// nearest (lowest face index on ties, like the ascending strict-'<' walk).  No plane is written.  Work ~ sum of bounding
// boxes instead of pixels x candidate faces: 3x faster than k_raster_fwd3<2> at 32 x 2048^2 (profiles/r02_*).
// ---------------------------------------------------------------------------------------------
#ifndef UMR_VISFACES_DYNAMIC
#define UMR_VISFACES_DYNAMIC 1   // warps fetch faces from a shared counter: same-box A/B 1.18 -> 0.97 ms at 32 x 2048^2 (0: static round-robin)
#endif
__global__ void __launch_bounds__(CTA) k_visible_faces(const float* __restrict__ rec_all, const uint16_t* __restrict__ clist,
                                                       const int* __restrict__ ccount, uint8_t* __restrict__ vis, Consts K) {
    __shared__ unsigned long long s_z[CB * CB];   // 32 KB
    __shared__ float s_xp[CB], s_yp[CB];
    __shared__ int s_bg, s_next;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.z, S = K.S, F = K.F;
    const int x0 = blockIdx.x * CB, y0 = blockIdx.y * CB;
    const int ncol = min(CB, S - x0), nrow = min(CB, S - y0);
    const size_t cidx = ((size_t)b * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    const int nc = __ldg(ccount + cidx);
    uint8_t* vb = vis + (size_t)b * F;
    if (nc == 0) {  // all background: face id -1, which the reference's indexing turns into face F-1 (see k_visible)
        if (tid == 0 && *reinterpret_cast<volatile uint8_t*>(vb + F - 1) == 0) vb[F - 1] = 1;
        return;
    }
    for (int i = tid; i < CB * CB; i += CTA) s_z[i] = ~0ull;
    if (tid < CB) s_xp[tid] = pixel_coord(x0 + tid, S);
    else if (tid < 2 * CB) s_yp[tid - CB] = pixel_coord(S - 1 - (y0 + tid - CB), S);
    if (tid == 0) { s_bg = 0; s_next = NWARP; }
    __syncthreads();
    const uint16_t* cl = clist + cidx * F;
    const float* rec_img = rec_all + (size_t)b * F * REC_F;
    auto scan_face = [&](int i) {
        const int f = __ldg(cl + i);
        const float* rc = rec_img + (size_t)f * REC_F;
        if (lane == 0 && i + NWARP < nc)   // the warp's next face: pull its 128-byte record into L1 while this one is scanned
            asm volatile("prefetch.global.L1 [%0];" ::"l"(rec_img + (size_t)__ldg(cl + i + NWARP) * REC_F));
        const uint32_t flg = __float_as_uint(__ldg(rc + R_FLG));
        if (!(K.double_side || (flg & 8u))) return;   // back face of a single-sided render never wins (warp-uniform)
        // pixel rectangle to test: bounding box of the vertices widened by 2 pixels (an inside pixel lies in the box; the
        // margin covers the rounding of the index conversion), or the whole cull box for a thin face
        float xlo, xhi, ylo, yhi;
        if (flg & 16u) {
            xlo = __ldg(rc + R_BOX); xhi = __ldg(rc + R_BOX + 1); ylo = __ldg(rc + R_BOX + 2); yhi = __ldg(rc + R_BOX + 3);
        } else {
            const float ax = __ldg(rc + 0), ay = __ldg(rc + 1), bx = __ldg(rc + 3), by = __ldg(rc + 4), cx = __ldg(rc + 6), cy = __ldg(rc + 7);
            xlo = fminf(fminf(ax, bx), cx); xhi = fmaxf(fmaxf(ax, bx), cx);
            ylo = fminf(fminf(ay, by), cy); yhi = fmaxf(fmaxf(ay, by), cy);
        }
        // pixel_coord(i) = (2 i + 1 - S) / S  <=>  i = (x S + S - 1) / 2;  rows run top-down: y index j = S - 1 - row
        int c0 = (int)floorf((xlo * (float)S + (float)(S - 1)) * 0.5f) - 2 - x0;
        int c1 = (int)ceilf((xhi * (float)S + (float)(S - 1)) * 0.5f) + 2 - x0;
        const int j0 = (int)floorf((ylo * (float)S + (float)(S - 1)) * 0.5f) - 2;
        const int j1 = (int)ceilf((yhi * (float)S + (float)(S - 1)) * 0.5f) + 2;
        int r0 = (S - 1 - j1) - y0, r1 = (S - 1 - j0) - y0;
        if (!(xlo == xlo && xhi == xhi && ylo == ylo && yhi == yhi)) { c0 = 0; c1 = CB; r0 = 0; r1 = CB; }  // NaN face: whole bin
        c0 = max(c0, 0); c1 = min(c1, ncol - 1); r0 = max(r0, 0); r1 = min(r1, nrow - 1);
        const int w = c1 - c0 + 1, h = r1 - r0 + 1;
        if (w <= 0 || h <= 0) return;   // warp-uniform
        const float i00 = __ldg(rc + R_INV + 0), i01 = __ldg(rc + R_INV + 1), i02 = __ldg(rc + R_INV + 2);
        const float i10 = __ldg(rc + R_INV + 3), i11 = __ldg(rc + R_INV + 4), i12 = __ldg(rc + R_INV + 5);
        const float i20 = __ldg(rc + R_INV + 6), i21 = __ldg(rc + R_INV + 7), i22 = __ldg(rc + R_INV + 8);
        const float4 bb = __ldg(reinterpret_cast<const float4*>(rc + R_BOX));
        const int n = w * h;
        for (int p = lane; p < n; p += 32) {
            const int lr = p / w, col = c0 + (p - lr * w), row = r0 + lr;
            const float xp = s_xp[col], yp = s_yp[row];
            if (xp > bb.y || xp < bb.x || yp > bb.w || yp < bb.z) continue;   // the cull test of the full kernel (kernel.cu:32-38)
            const float w0 = i00 * xp + i01 * yp + i02;
            const float w1 = i10 * xp + i11 * yp + i12;
            const float w2 = i20 * xp + i21 * yp + i22;
            if (!(w0 <= 1 && w0 >= 0 && w1 <= 1 && w1 >= 0 && w2 <= 1 && w2 >= 0)) continue;   // kernel.cu:404
            bool pass = w0 > 0 && w1 > 0 && w2 > 0 && w0 < 1 && w1 < 1 && w2 < 1;
            if (!pass) {  // a barycentric exactly 0 or 1: the reference's outside branch decides (kernel.cu:380-383)
                Frag fr;
                pass = fragment(rc, xp, yp, K.thr, K.sigma, fr);
            }
            if (!pass) continue;
            float k0 = w0, k1 = w1, k2 = w2;
            clip_bary(k0, k1, k2);
            const float zp = depth_of(rc, k0, k1, k2);
            if (zp < K.near_ || zp > K.far_ || !(zp < 10000000.f)) continue;   // depth range (:399), initial depth_min (:341)
            // zp > 0 here (near_ >= 0 is assumed by the packed ordering; negative depths are ordered by the sign fix below)
            uint32_t zb = __float_as_uint(zp);
            zb = (zb & 0x80000000u) ? ~zb : (zb | 0x80000000u);   // total order of floats as unsigned integers
            atomicMin(&s_z[row * CB + col], ((unsigned long long)zb << 32) | (unsigned long long)(uint32_t)f);
        }
    };
#if UMR_VISFACES_DYNAMIC
    // faces are handed out dynamically (shared counter): their bounding boxes differ by an order of magnitude and a static
    // round-robin left 27 % of the stall samples at the end-of-bin barrier (profiles/r02_C3_step_ncu_summary.txt)
    for (int i = warp; i < nc;) {
        scan_face(i);
        int nx = 0;
        if (lane == 0) nx = atomicAdd(&s_next, 1);
        i = __shfl_sync(0xffffffffu, nx, 0);
    }
#else
    for (int i = warp; i < nc; i += NWARP) scan_face(i);
#endif
    __syncthreads();
    bool bg = false;
    for (int i = tid; i < CB * CB; i += CTA) {
        const int row = i / CB, col = i - row * CB;
        if (row >= nrow || col >= ncol) continue;
        const unsigned long long key = s_z[i];
        if (key == ~0ull) { bg = true; continue; }
        uint8_t* m = vb + (uint32_t)(key & 0xffffffffull);
        if (*reinterpret_cast<volatile uint8_t*>(m) == 0) *m = 1;
    }
    if (bg) s_bg = 1;   // benign race: every writer stores 1
    __syncthreads();
    if (tid == 0 && s_bg && *reinterpret_cast<volatile uint8_t*>(vb + F - 1) == 0) vb[F - 1] = 1;
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
template <int RGB>
__global__ void __launch_bounds__(CTA, UMR_FWD2_CTAS) k_raster_fwd2(const float* __restrict__ rec_all, const float4* __restrict__ box_all,
                                                        const uint16_t* __restrict__ clist, const int* __restrict__ ccount,
                                                        const float* __restrict__ textures, float* __restrict__ images,
                                                        float* __restrict__ colors_hi, float* __restrict__ aggrs,
                                                        float* __restrict__ p2f_acc, const uint32_t* __restrict__ ubox,
                                                        Consts K, float eps, float bg0, float bg1, float bg2, PairBuf pb,
                                                        int ncb) {
    // dynamic shared memory: two slot buffers (phase A of sub-chunk i+1 overlaps phase B of sub-chunk i)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* s_sD = reinterpret_cast<float*>(smem_raw);                       // [2][SLOTS]
    float* s_sZ = s_sD + 2 * SLOTS;                                         // [2][SLOTS]
    uint32_t* s_sT = reinterpret_cast<uint32_t*>(s_sZ + 2 * SLOTS);         // [2][SLOTS]
    __shared__ __align__(128) float s_rec[NSTAGE * CHUNK * REC_F];  // 8 KB; reused by the store epilogue
    __shared__ uint16_t s_list[LCAP];
    __shared__ uint32_t s_geo[LCAP];
    __shared__ uint32_t s_boff[LCAP + 1];
    __shared__ float s_xp[TILE], s_yp[TILE], s_ext[4];
    __shared__ int s_warp_cnt[NWARP];
    __shared__ uint32_t s_warp_blk[NWARP];
    __shared__ uint32_t s_segbase;  // first record block of the current segment (valid while s_save)
    __shared__ int s_save;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.z;
    const int S = K.S, F = K.F;
    const int tx0 = blockIdx.x * TILE, ty0 = blockIdx.y * TILE;
    const size_t np = (size_t)S * S;

    tile_extents(S, s_ext);
    if (tid < TILE) s_xp[tid] = pixel_coord(tx0 + tid, S);
    else if (tid < 2 * TILE) s_yp[tid - TILE] = pixel_coord(S - 1 - (ty0 + tid - TILE), S);
    if (tid == 0) s_save = pb.cap > 0 ? 1 : 0;
    __syncthreads();

    const size_t tile_id = ((size_t)b * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    const size_t cidx = ((size_t)b * ncb + (ty0 / CB)) * ncb + (tx0 / CB);
    const int nc = tile_outside_union(ubox, b, s_ext) ? 0 : __ldg(ccount + cidx);

    if (nc == 0) {
        // ---- untouched tile (most of the image): every pixel holds the initial state.  Same arithmetic as the
        // general path (kernel.cu:335-348, 443-475), evaluated once, stored with 128-bit stores where possible.
        if (tid == 0 && pb.cap > 0) pb.tile_head[tile_id] = TILE_EMPTY;
        const float ssum0 = expf(eps / K.gamma);
        float o0, o1, o2, g0, g1;
        if (RGB == 0) {
            o0 = bg0; o1 = bg1; o2 = bg2;
            g0 = 10000000.f; g1 = -1.f;
        } else {
            const float q0 = bg0 * ssum0, q1 = bg1 * ssum0, q2 = bg2 * ssum0;
            o0 = q0 == 0.f ? q0 : q0 / ssum0;
            o1 = q1 == 0.f ? q1 : q1 / ssum0;
            o2 = q2 == 0.f ? q2 : q2 / ssum0;
            g0 = ssum0; g1 = eps;
        }
        const float alpha = (float)(1. - (double)1.f);
        const float full[6] = {o0, o1, o2, alpha, g0, g1};
        float pooled[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) pooled[k] = (((full[k] + full[k]) + full[k]) + full[k]) * 0.25f;
        if (K.aa && K.vec_store && tx0 + TILE <= S && ty0 + TILE <= S) {
            for (int i = tid; i < 6 * 64; i += CTA) {
                const int plane = i >> 6, rem = i & 63, row = rem >> 2, q = rem & 3;
                float x = full[0];
#pragma unroll
                for (int k = 1; k < 6; ++k) x = (plane == k) ? full[k] : x;
                const float4 val = make_float4(x, x, x, x);
                const size_t off = (size_t)(ty0 + row) * S + tx0 + q * 4;
                if (plane < 4) {
                    if (colors_hi != nullptr)
                        *reinterpret_cast<float4*>(colors_hi + ((size_t)b * 4 + plane) * np + off) = val;
                } else {
                    *reinterpret_cast<float4*>(aggrs + ((size_t)b * 2 + (plane - 4)) * np + off) = val;
                }
            }
            if (tid < 64) {
                const int k = tid >> 4, rem = tid & 15, row = rem >> 1, q = rem & 1;
                float x = pooled[0];
#pragma unroll
                for (int kk = 1; kk < 4; ++kk) x = (k == kk) ? pooled[kk] : x;
                const int IS = K.IS;
                const size_t nq = (size_t)IS * IS;
                *reinterpret_cast<float4*>(images + ((size_t)b * 4 + k) * nq + (size_t)((ty0 >> 1) + row) * IS + (tx0 >> 1) + q * 4) =
                    make_float4(x, x, x, x);
            }
            return;
        }
        const int px = tx0 + (tid & (TILE - 1)), py = ty0 + (tid >> 4);
        if (px < S && py < S) {
            const size_t p = (size_t)py * S + px;
            aggrs[((size_t)b * 2 + 0) * np + p] = g0;
            aggrs[((size_t)b * 2 + 1) * np + p] = g1;
            if (colors_hi != nullptr) {
#pragma unroll
                for (int k = 0; k < 4; ++k) colors_hi[((size_t)b * 4 + k) * np + p] = full[k];
            }
            if (K.aa) {
                if ((px & 1) == 0 && (py & 1) == 0 && px + 1 < S && py + 1 < S) {
                    const size_t q = (size_t)(py >> 1) * K.IS + (px >> 1);
                    const size_t nq = (size_t)K.IS * K.IS;
#pragma unroll
                    for (int k = 0; k < 4; ++k) images[((size_t)b * 4 + k) * nq + q] = pooled[k];
                }
            } else if (images != colors_hi) {
#pragma unroll
                for (int k = 0; k < 4; ++k) images[((size_t)b * 4 + k) * np + p] = full[k];
            }
        }
        return;
    }

    const PixelMap pm = map_pixel(S);
    const int px = pm.px, py = pm.py;
    const bool live = pm.live;
    const int lcol = px - tx0, lrow = py - ty0;       // pixel position inside the tile
    const int ncol = min(TILE, S - tx0), nrow = min(TILE, S - ty0);
    const int bx0 = (warp & 1) * 8, by0 = (warp >> 1) * 4;  // this warp's 8x4 pixel block inside the tile
    const uint16_t* cl = clist + cidx * F;
    const float4* box = box_all + (size_t)b * F;
    const float* rec_img = rec_all + (size_t)b * F * REC_F;
    const float* tex_img = textures + (size_t)(b / K.tex_div) * K.tex_bs;
    const float ext0 = s_ext[0], ext1 = s_ext[1], ext2 = s_ext[2], ext3 = s_ext[3];

    // pixel state (kernel.cu:335-348)
    float acc_a = 1.f;
    float ssum = expf(eps / K.gamma);
    float smax = eps;
    float c0, c1, c2;
    if (RGB == 1) { c0 = bg0 * ssum; c1 = bg1 * ssum; c2 = bg2 * ssum; }
    else { c0 = bg0; c1 = bg1; c2 = bg2; }
    float zmin = 10000000.f;
    int fid = -1;
    // torch-1.1 affine_grid (align_corners=True) coordinates of this pixel: linspace(-1, 1, S)
    const float gstep = 2.f / (float)(S - 1);
    const float gx = (px * 2 < S) ? (-1.f + gstep * px) : (1.f - gstep * (S - 1 - px));
    const float gy = (py * 2 < S) ? (-1.f + gstep * py) : (1.f - gstep * (S - 1 - py));

    int32_t head = TILE_EMPTY;     // meaningful in thread 0
    uint32_t prev_seg = SEG_NONE;  // meaningful in thread 0
    const uint32_t lt = (1u << lane) - 1u;

    for (int w0 = 0; w0 < nc; w0 += LCAP) {
        // ---- tile-list segment: ordered compaction of this window's coarse entries that touch the tile ------
        const int nwin = min(LCAP, nc - w0);
        uint32_t masks[LCAP / CTA], geos[LCAP / CTA];
        uint16_t fids[LCAP / CTA];
        int cnt = 0;
#pragma unroll
        for (int r = 0; r < LCAP / CTA; ++r) {
            const int i = warp * (LCAP / NWARP) + r * 32 + lane;
            bool hit = false;
            uint32_t geo = 0;
            uint16_t f = 0;
            if (i < nwin) {
                f = __ldg(cl + w0 + i);
                const float4 bb = __ldg(box + f);
                hit = !(ext0 > bb.y || ext1 < bb.x || ext2 > bb.w || ext3 < bb.z);
                if (hit) {
                    // rectangle of tile pixels passing the per-pixel cull test !(xp > hi || xp < lo || ...) (kernel.cu:32-38);
                    // pixel-centre coordinates are monotone, so it is [c_lo, c_hi) x [r_lo, r_hi).  Counting with the SAME
                    // comparisons keeps NaN boxes "never culled", like the per-pixel form.
                    int c_lo = 0, c_gt = 0, r_lo = 0, r_lt = 0;
#pragma unroll
                    for (int q = 0; q < TILE; ++q) {
                        const float x = s_xp[q], y = s_yp[q];
                        c_lo += (q < ncol && x < bb.x) ? 1 : 0;
                        c_gt += (q < ncol && x > bb.y) ? 1 : 0;
                        r_lo += (q < nrow && y > bb.w) ? 1 : 0;
                        r_lt += (q < nrow && y < bb.z) ? 1 : 0;
                    }
                    const int w = max(0, ncol - c_gt - c_lo), h = max(0, nrow - r_lt - r_lo);
                    geo = (uint32_t)c_lo | ((uint32_t)w << 4) | ((uint32_t)r_lo << 9) | ((uint32_t)h << 13);
                }
            }
            masks[r] = __ballot_sync(0xffffffffu, hit);
            geos[r] = geo;
            fids[r] = f;
            cnt += __popc(masks[r]);
        }
        if (lane == 0) s_warp_cnt[warp] = cnt;
        __syncthreads();
        int off = 0, n = 0;
#pragma unroll
        for (int w = 0; w < NWARP; ++w) {
            const int c = s_warp_cnt[w];
            if (w < warp) off += c;
            n += c;
        }
#pragma unroll
        for (int r = 0; r < LCAP / CTA; ++r) {
            if ((masks[r] >> lane) & 1u) {
                const int pos = off + __popc(masks[r] & lt);
                s_list[pos] = fids[r];
                s_geo[pos] = geos[r];
            }
            off += __popc(masks[r]);
        }
        __syncthreads();  // list + geo visible; s_warp_cnt reusable
        if (n == 0) continue;  // uniform

        // ---- block offsets: exclusive prefix of ceil(w*h / 32) over the segment (2 entries per thread) ---------
        {
            uint32_t v0 = 0, v1 = 0;
            const int i0 = 2 * tid, i1 = 2 * tid + 1;
            if (i0 < n) { const uint32_t g = s_geo[i0]; v0 = (((g >> 4) & 31u) * ((g >> 13) & 31u) + 31u) >> 5; }
            if (i1 < n) { const uint32_t g = s_geo[i1]; v1 = (((g >> 4) & 31u) * ((g >> 13) & 31u) + 31u) >> 5; }
            uint32_t incl = v0 + v1;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += o;
            }
            if (lane == 31) s_warp_blk[warp] = incl;
            __syncthreads();
            uint32_t base = 0, total = 0;
#pragma unroll
            for (int w = 0; w < NWARP; ++w) {
                const uint32_t c = s_warp_blk[w];
                base += (w < warp) ? c : 0u;
                total += c;
            }
            const uint32_t excl = base + incl - (v0 + v1);
            if (i0 < n) s_boff[i0] = excl;
            if (i1 < n) s_boff[i1] = excl + v0;
            if (tid == 0) s_boff[n] = total;
        }
        __syncthreads();
        const uint32_t NBw = s_boff[n];
        if (NBw == 0) continue;  // uniform: every rectangle is empty

        // ---- reserve the segment's blocks in the pair buffer --------------------------------------------------
        if (tid == 0 && *reinterpret_cast<volatile int*>(&s_save)) {  // thread 0 only (volatile: the read is not hoisted)
            const uint32_t base = atomicAdd(pb.ctrl, NBw + 2u);
            if ((uint64_t)base + NBw + 2u > (uint64_t)pb.cap) {
                s_save = 0;  // does not fit: the whole tile falls back to the recompute backward
                head = TILE_UNSAVED;
                pb.ulist[atomicAdd(pb.ctrl + 1, 1u)] = (int32_t)tile_id;
            } else {
                pb.blk_hdr[base] = NBw;
                pb.blk_hdr[base + 1] = SEG_NONE;
                if (prev_seg == SEG_NONE) head = (int32_t)base;
                else pb.blk_hdr[prev_seg + 1] = base;
                prev_seg = base;
                s_segbase = base + 2u;
            }
        }

        issue_chunk(rec_img, s_list, n, 0, s_rec);
        issue_chunk(rec_img, s_list, n, 1, s_rec);
        cp_async_wait<1>();
        __syncthreads();  // chunk 0 landed for every thread; s_save / s_segbase visible
        const bool save = s_save != 0;
        const uint32_t segbase = s_segbase;

        // ---------------- phase A: pair-parallel geometry of faces [ja, jb) (blocks [kb0, kb1)) -> slot buffer `buf`
        auto phase_a = [&](int c, int ja, uint32_t kb0, uint32_t kb1, int buf) {
            const float* chunk = s_rec + (size_t)(c % NSTAGE) * CHUNK * REC_F;
            const int cbeg = c * CHUNK;
            float* sD = s_sD + buf * SLOTS;
            float* sZ = s_sZ + buf * SLOTS;
            uint32_t* sT = s_sT + buf * SLOTS;
            // each warp owns a contiguous range of the sub-chunk's blocks: the face index only moves forward and the
            // face's rectangle is decoded once per face, not once per block
            const uint32_t nblk = kb1 - kb0, per = (nblk + NWARP - 1) / NWARP;
            const uint32_t kbeg = kb0 + min(nblk, warp * per), kend = kb0 + min(nblk, (warp + 1) * per);
            int j = ja;
            uint32_t jnext = s_boff[ja + 1], jbase = kb0;  // blocks of face j: [jbase, jnext)
            int cx0 = 0, w = 0, ry0 = 0, size = 0;
            uint32_t rcpw = 0;
            bool fresh = true;
            for (uint32_t k = kbeg; k < kend; ++k) {
                while (jnext <= k) { ++j; jnext = s_boff[j + 1]; fresh = true; }  // warp-uniform, monotone
                if (fresh) {
                    const uint32_t geo = s_geo[j];
                    cx0 = (int)(geo & 15u); w = (int)((geo >> 4) & 31u);
                    ry0 = (int)((geo >> 9) & 15u);
                    size = w * (int)((geo >> 13) & 31u);
                    rcpw = c_rcpw[w];
                    jbase = s_boff[j];
                    fresh = false;
                }
                const int local = (int)(k - jbase) * 32 + lane;
                const int slot = (int)(k - kb0) * 32 + lane;
                const float* rc = chunk + (j - cbeg) * REC_F;
                uint32_t tflags = 0;
                bool emit = false;
                Frag fr;
                float k0 = 0.f, k1 = 0.f, k2 = 0.f, zp = 0.f;
                int pix = 0;
                if (local < size) {
                    const int lr = (int)(((uint32_t)local * rcpw) >> 16);  // exact floor(local / w): local < 1024, w <= 32
                    const int col = cx0 + (local - lr * w), row = ry0 + lr;
                    pix = row * TILE + col;
                    if (fragment(rc, s_xp[col], s_yp[row], K.thr, K.sigma, fr)) {
                        k0 = fr.w0; k1 = fr.w1; k2 = fr.w2;
                        clip_bary(k0, k1, k2);
                        zp = depth_of(rc, k0, k1, k2);
                        const bool zv = !(zp < K.near_ || zp > K.far_);
                        const uint32_t flg = __float_as_uint(rc[R_FLG]);
                        tflags = SL_VALID | (zv ? SL_ZV : 0u) | ((flg & 8u) ? SL_FRONT : 0u) |
                                 (uint32_t)texel_index(k0, k1, K.R);
                        if (RGB == 0) {
                            const bool inside = fr.w0 <= 1 && fr.w0 >= 0 && fr.w1 <= 1 && fr.w1 >= 0 &&
                                                fr.w2 <= 1 && fr.w2 >= 0;
                            if (inside) tflags |= SL_INS;
                        } else {
                            // normalised depth (kernel.cu:418).  The backward needs THESE bits: its softmax weight is
                            // exp((zn - max) / gamma), and a 1-ulp change of zn is a 5e-4 relative change of the weight.
                            zp = (K.far_ - zp) / (K.far_ - K.near_);
                        }
                        sZ[slot] = zp;
                        sD[slot] = fr.D;
                        emit = zv;  // kernel.cu:592 drops every gradient of an out-of-range pair
                    }
                    sT[slot] = tflags;
                }
                if (save) {  // uniform
                    const uint32_t m = __ballot_sync(0xffffffffu, emit);
                    if (emit) {
                        const int pos = __popc(m & lt);
                        float4* dst = pb.recs + (size_t)(segbase + k) * BLK_F4 + pos;
                        // closest-point barycentrics as the reference forms them: t_k + w_k (kernel.cu:638-641)
                        const float u0 = fr.t0 + fr.w0, u1 = fr.t1 + fr.w1, u2 = fr.t2 + fr.w2;
                        const uint32_t meta = (uint32_t)pix | ((tflags & SL_TIX) << 8) | ((tflags & SL_FRONT) ? (1u << 24) : 0u);
                        dst[0] = make_float4(fr.D, fr.sign * fr.dx, fr.sign * fr.dy, zp);  // zp: depth (hard) / normalised depth (softmax)
                        dst[32] = make_float4(u0, u1, u2, __uint_as_float(meta));
                        // w_clip_k / z_k^2 (kernel.cu:624-627) through the record's precomputed 1 / z_k^2
                        dst[64] = make_float4(k0 * rc[R_IZ2], k1 * rc[R_IZ2 + 1], k2 * rc[R_IZ2 + 2], 0.f);
                    }
                    if (lane == 0) pb.blk_hdr[segbase + k] = (uint32_t)s_list[j] | ((uint32_t)__popc(m) << 16);
                }
            }
        };

        float own_x = 0.f, own_y = 0.f, own_w = 0.f;  // p2f partial sums: lane j owns face j of the current chunk

        // ---------------- phase B: ordered per-pixel aggregation of faces [ja, jb) from slot buffer `buf`
        auto phase_b = [&](int c, int ja, int jb, uint32_t kb0, int buf) {
            const int cbeg = c * CHUNK;
            const float* sD = s_sD + buf * SLOTS;
            const float* sZ = s_sZ + buf * SLOTS;
            const uint32_t* sT = s_sT + buf * SLOTS;
            // faces of the sub-chunk whose rectangle meets this warp's 8x4 block (lane i <-> face ja + i)
            bool meets = false;
            if (ja + lane < jb) {
                const uint32_t g = s_geo[ja + lane];
                const int cx0 = (int)(g & 15u), w = (int)((g >> 4) & 31u);
                const int ry0 = (int)((g >> 9) & 15u), h = (int)((g >> 13) & 31u);
                meets = w > 0 && h > 0 && cx0 < bx0 + 8 && cx0 + w > bx0 && ry0 < by0 + 4 && ry0 + h > by0;
            }
            uint32_t fm = __ballot_sync(0xffffffffu, meets);
            while (fm) {
                const int i = __ffs(fm) - 1;
                fm &= fm - 1u;
                const int j = ja + i;
                const uint32_t geo = s_geo[j];
                const int cx0 = (int)(geo & 15u), w = (int)((geo >> 4) & 31u);
                const int ry0 = (int)((geo >> 9) & 15u), h = (int)((geo >> 13) & 31u);
                const int dc = lcol - cx0, dr = lrow - ry0;
                float a_x = 0.f, a_y = 0.f, a_w = 0.f;
                bool contrib = false;
                if (live && (unsigned)dc < (unsigned)w && (unsigned)dr < (unsigned)h) {
                    const int slot = (int)(s_boff[j] - kb0) * 32 + dr * w + dc;
                    const uint32_t t = sT[slot];
                    if (t & SL_VALID) {
                        const float D = sD[slot];
                        acc_a = (float)((double)acc_a * (1. - (double)D));  // kernel.cu:396
                        if (t & SL_ZV) {
                            const int f = s_list[j];
                            const bool front = (t & SL_FRONT) != 0;
                            if (RGB == 0) {
                                const float zp = sZ[slot];
                                if (zp < zmin && (t & SL_INS) && (K.double_side || front)) {
                                    zmin = zp;
                                    fid = f;
                                    const float* tp = tex_img + ((size_t)f * K.T2 + (t & SL_TIX)) * 3;
                                    c0 = __ldg(tp); c1 = __ldg(tp + 1); c2 = __ldg(tp + 2);
                                }
                            } else if (front || K.double_side) {
                                const float zn = sZ[slot];
                                float ed = 1.f;
                                if (zn > smax) { ed = expf((smax - zn) / K.gamma); smax = zn; }
                                const float ez = expf((zn - smax) / K.gamma);
                                ssum = ed * ssum + ez * D;
                                const float a = ez * D;
                                if (a != 0.f || ed != 1.f) {  // else: c = 1*c + 0*texel, p2f terms 0 (exact)
                                    a_x = a * gx; a_y = a * gy; a_w = a;
                                    contrib = a != 0.f;
                                    const float* tp = tex_img + ((size_t)f * K.T2 + (t & SL_TIX)) * 3;
                                    c0 = ed * c0 + a * __ldg(tp);
                                    c1 = ed * c1 + a * __ldg(tp + 1);
                                    c2 = ed * c2 + a * __ldg(tp + 2);
                                }
                            }
                        }
                    }
                }
                if (RGB == 1 && p2f_acc != nullptr) {
                    if (__any_sync(0xffffffffu, contrib)) {
                        a_x = warp_sum(a_x); a_y = warp_sum(a_y); a_w = warp_sum(a_w);
                        if (lane == j - cbeg) { own_x += a_x; own_y += a_y; own_w += a_w; }
                    }
                }
            }
        };
        // last face (exclusive) of the sub-chunk starting at ja inside chunk c: at most SUB_BLOCKS blocks
        auto sub_end = [&](int c, int ja) {
            const int cend = min(n, (c + 1) * CHUNK);
            const uint32_t kb0 = s_boff[ja];
            int jb = ja + 1;
            while (jb < cend && s_boff[jb + 1] - kb0 <= (uint32_t)SUB_BLOCKS) ++jb;
            return jb;
        };

        // Software pipeline over the segment's sub-chunks: A(0) | bar | B(0) A(1) | bar | B(1) A(2) | ... so the
        // (pixel-imbalanced) aggregation of one sub-chunk overlaps the (balanced) geometry of the next and there is ONE
        // CTA barrier per sub-chunk.  Record stage c % 2 is refilled once every phase A on chunk c is complete.
        int c = 0, ja = 0, jb = sub_end(0, 0), buf = 0;
        while (true) {
            phase_a(c, ja, s_boff[ja], s_boff[jb], buf);
            const int cend = min(n, (c + 1) * CHUNK);
            const bool has_next = jb < n;
            const bool chunk_switch = has_next && jb >= cend;  // the next sub-chunk starts chunk c + 1
            if (chunk_switch) cp_async_wait<0>();               // ... whose records (the only group in flight) have landed
            __syncthreads();  // A(cur) complete; B(prev) complete (its slot buffer is free); chunk c + 1 visible
            if (chunk_switch) issue_chunk(rec_img, s_list, n, c + NSTAGE, s_rec);  // stage c % 2 is free now
            phase_b(c, ja, jb, s_boff[ja], buf);
            if (jb >= cend) {  // chunk c fully aggregated: one global RED per (warp, face, component)
                if (RGB == 1 && p2f_acc != nullptr && own_w != 0.f) {
                    float* dst = p2f_acc + ((size_t)b * F + s_list[c * CHUNK + lane]) * 4;
                    red_add_global(dst + 0, own_x);
                    red_add_global(dst + 1, own_y);
                    red_add_global(dst + 2, own_w);
                }
                own_x = own_y = own_w = 0.f;
            }
            if (!has_next) break;
            if (chunk_switch) ++c;
            ja = jb;
            jb = sub_end(c, ja);
            buf ^= 1;
        }
        cp_async_wait<0>();
        __syncthreads();  // segment done: s_list / s_geo / s_boff / s_rec / slots reusable
    }
    if (tid == 0 && pb.cap > 0) pb.tile_head[tile_id] = head;

    // ---- finalise (kernel.cu:443-475) + fused 2x2 pool + coalesced stores (as round 1) --------------------
    const float alpha = (float)(1. - (double)acc_a);  // kernel.cu:449-451
    float o0, o1, o2, g0, g1;
    if (RGB == 0) {
        o0 = c0; o1 = c1; o2 = c2;
        g0 = zmin; g1 = (float)fid;
    } else {
        o0 = c0 == 0.f ? c0 : c0 / ssum;
        o1 = c1 == 0.f ? c1 : c1 / ssum;
        o2 = c2 == 0.f ? c2 : c2 / ssum;
        g0 = ssum; g1 = smax;
    }
    float v[4] = {o0, o1, o2, alpha};
    if (K.aa) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float a01 = __shfl_xor_sync(0xffffffffu, v[k], 1);
            const float a10 = __shfl_xor_sync(0xffffffffu, v[k], 8);
            const float a11 = __shfl_xor_sync(0xffffffffu, v[k], 9);
            v[k] = (((v[k] + a01) + a10) + a11) * 0.25f;  // meaningful on the (even x, even y) lane
        }
    }
    if (K.aa && K.vec_store && tx0 + TILE <= S && ty0 + TILE <= S) {  // uniform: full tile, aligned buffers
        float* st = s_rec;  // 6 * 256 + 4 * 64 = 1792 floats <= 2048
        const int o = lrow * TILE + lcol;
        st[0 * 256 + o] = o0; st[1 * 256 + o] = o1; st[2 * 256 + o] = o2; st[3 * 256 + o] = alpha;
        st[4 * 256 + o] = g0; st[5 * 256 + o] = g1;
        if ((lane & 1) == 0 && (lane & 8) == 0) {
            const int po = (lrow >> 1) * (TILE / 2) + (lcol >> 1);
#pragma unroll
            for (int k = 0; k < 4; ++k) st[6 * 256 + k * 64 + po] = v[k];
        }
        __syncthreads();
        for (int i = tid; i < 6 * 64; i += CTA) {
            const int plane = i >> 6, rem = i & 63, row = rem >> 2, q = rem & 3;
            const float4 val = *reinterpret_cast<const float4*>(st + plane * 256 + row * TILE + q * 4);
            const size_t off = (size_t)(ty0 + row) * S + tx0 + q * 4;
            if (plane < 4) {
                if (colors_hi != nullptr)
                    *reinterpret_cast<float4*>(colors_hi + ((size_t)b * 4 + plane) * np + off) = val;
            } else {
                *reinterpret_cast<float4*>(aggrs + ((size_t)b * 2 + (plane - 4)) * np + off) = val;
            }
        }
        if (tid < 64) {
            const int k = tid >> 4, rem = tid & 15, row = rem >> 1, q = rem & 1;
            const float4 val = *reinterpret_cast<const float4*>(st + 6 * 256 + k * 64 + row * (TILE / 2) + q * 4);
            const int IS = K.IS;
            const size_t nq = (size_t)IS * IS;
            *reinterpret_cast<float4*>(images + ((size_t)b * 4 + k) * nq + (size_t)((ty0 >> 1) + row) * IS + (tx0 >> 1) + q * 4) = val;
        }
        return;
    }
    if (live) {
        const size_t p = (size_t)py * S + px;
        aggrs[((size_t)b * 2 + 0) * np + p] = g0;
        aggrs[((size_t)b * 2 + 1) * np + p] = g1;
        if (colors_hi != nullptr) {
            colors_hi[((size_t)b * 4 + 0) * np + p] = o0;
            colors_hi[((size_t)b * 4 + 1) * np + p] = o1;
            colors_hi[((size_t)b * 4 + 2) * np + p] = o2;
            colors_hi[((size_t)b * 4 + 3) * np + p] = alpha;
        }
    }
    if (K.aa) {
        if (live && (lane & 1) == 0 && (lane & 8) == 0) {
            const int IS = K.IS;
            const size_t q = (size_t)(py >> 1) * IS + (px >> 1);
            const size_t nq = (size_t)IS * IS;
#pragma unroll
            for (int k = 0; k < 4; ++k) images[((size_t)b * 4 + k) * nq + q] = v[k];
        }
    } else if (live && images != colors_hi) {
        const size_t p = (size_t)py * S + px;
        images[((size_t)b * 4 + 0) * np + p] = o0;
        images[((size_t)b * 4 + 1) * np + p] = o1;
        images[((size_t)b * 4 + 2) * np + p] = o2;
        images[((size_t)b * 4 + 3) * np + p] = alpha;
    }
}

// ---------------------------------------------------------------------------------------------
// backward: stream the saved pair records of the tile
// ---------------------------------------------------------------------------------------------
// NC consecutive texel-gradient floats: RGB goes out as one 2-float vector RED + one scalar (common.cuh)
template <int NC>
__device__ __forceinline__ void red_add_texel(float* gt, const float* v) {
    if (NC == 3) {
        red_add3_global(gt, v[0], v[1], v[2]);
    } else {
#pragma unroll
        for (int c = 0; c < NC; ++c) red_add_global(gt + c, v[c]);
    }
}

#ifndef UMR_BWD2_REGPIPE
#define UMR_BWD2_REGPIPE 0
#endif
#ifndef UMR_BWD2_CTAS
#define UMR_BWD2_CTAS 5   // same-box A/B (profiles/r02_bwd2_cta_ab.txt): 5 CTAs x 256 threads (48 registers) beats 4 x 256 by 5 % at C2 and
#endif                    // 4 % at 8 x 2048^2; 8-10 CTAs x 128 threads gain up to 8 % on the large render only; 64-thread CTAs lose

#ifndef UMR_BWD2_THREADS
#define UMR_BWD2_THREADS 256
#endif
constexpr int BWD2_THREADS = UMR_BWD2_THREADS, BWD2_WARPS = BWD2_THREADS / 32;  // threads of one k_raster_bwd2 CTA (one tile)
// TS: side of the forward's tile (16: k_raster_fwd3 / k_raster_fwd2, 32: k_raster_fwd4); one CTA streams one tile
// GEOM = false: the caller wants no gradient for the vertices (UMR's texture branch renders DETACHED geometry,
// experiments/train_s2.py:248) -- only the texel gradients are formed, the compiler drops the rest of the arithmetic.
// PRE = true: the texel gradients of a step are combined inside the warp before they go to global memory -- lanes that hit
// the same texel of the step's face (found with match.any) are summed by the lowest of them, which issues the only REDs.
// Pays only when a face covers hundreds of raster pixels per texel (raster.cu: UMR_TEXGRAD_PRE_RATIO_*); at UMR's shapes the
// plain vector REDs are faster (profiles/r02_texgrad_pre_ab2.txt).
template <int RGB, bool TEXGRAD, int TS, int NC = 3, bool GEOM = true, bool PRE = false>  // NC colour channels; pixel planes: g[NC], g_alpha, C[NC], alpha, ssum, smax
__global__ void __launch_bounds__(BWD2_THREADS, UMR_BWD2_CTAS) k_raster_bwd2(const float* __restrict__ textures, const float* __restrict__ colors_hi,
                                                        const float* __restrict__ aggrs, const float* __restrict__ grad_images,
                                                        float* __restrict__ grad_faces, float* __restrict__ grad_tex, Consts K,
                                                        PairBuf pb) {
    constexpr int NPL = NC + 1, NV = 2 * NPL + 2;
    __shared__ float s_pix[NV][TS * TS];  // g[NC], g_alpha, C[NC], alpha, ssum, smax (row-major tile pixels)
    constexpr int NP = TS * TS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.z;
    const int S = K.S, F = K.F;
    const size_t tile_id = ((size_t)b * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    const int32_t head = __ldg(pb.tile_head + tile_id);
    if (head < 0) return;  // empty, or unsaved (k_raster_bwd_pairs handles it)
    const int x0 = blockIdx.x * TS, y0 = blockIdx.y * TS;
    for (int pi = tid; pi < NP; pi += BWD2_THREADS) {
        const int px = x0 + (pi % TS), py = y0 + (pi / TS);
        const size_t np = (size_t)S * S;
        float v[NV];
#pragma unroll
        for (int k = 0; k < NV; ++k) v[k] = (k == NV - 2) ? 1.f : 0.f;
        if (px < S && py < S) {
            const size_t p = (size_t)py * S + px;
            if (K.aa) {  // avg_pool2d backward: g / 4
                const size_t nq = (size_t)K.IS * K.IS;
                const size_t q = (size_t)(py >> 1) * K.IS + (px >> 1);
#pragma unroll
                for (int k = 0; k < NPL; ++k) v[k] = __ldg(grad_images + ((size_t)b * NPL + k) * nq + q) * 0.25f;
            } else {
#pragma unroll
                for (int k = 0; k < NPL; ++k) v[k] = __ldg(grad_images + ((size_t)b * NPL + k) * np + p);
            }
#pragma unroll
            for (int k = 0; k < NPL; ++k) v[NPL + k] = __ldg(colors_hi + ((size_t)b * NPL + k) * np + p);
            v[NV - 2] = __ldg(aggrs + ((size_t)b * 2 + 0) * np + p);
            v[NV - 1] = __ldg(aggrs + ((size_t)b * 2 + 1) * np + p);
        }
#pragma unroll
        for (int k = 0; k < NV; ++k) s_pix[k][pi] = v[k];
    }
    __syncthreads();
    const float* tex_img = textures + (size_t)(b / K.tex_div) * K.tex_bs;
    float* gtex_img = TEXGRAD ? grad_tex + (size_t)(b / K.tex_div) * K.tex_bs : nullptr;
    float* gf_img = grad_faces + (size_t)b * F * 9;

    float acc[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[k] = 0.f;
    int cur_f = -1;
    auto flush = [&]() {
        if (GEOM && cur_f >= 0) {
#pragma unroll
            for (int k = 0; k < 9; ++k) acc[k] = warp_sum(acc[k]);
            if (lane < 9) {
                float v = acc[0];
#pragma unroll
                for (int k = 1; k < 9; ++k) v = (lane == k) ? acc[k] : v;
                if (v != 0.f) red_add_global(gf_img + (size_t)cur_f * 9 + lane, v);
            }
#pragma unroll
            for (int k = 0; k < 9; ++k) acc[k] = 0.f;
        }
    };

    uint32_t seg = (uint32_t)head;
    while (seg != SEG_NONE) {
        const uint32_t NB = __ldg(pb.blk_hdr + seg), next = __ldg(pb.blk_hdr + seg + 1);
        const uint32_t per = (NB + BWD2_WARPS - 1) / BWD2_WARPS;
        const uint32_t kbeg = min(NB, warp * per), kend = min(NB, kbeg + per);
        const uint32_t* hdrs = pb.blk_hdr + seg + 2;
        // The warp streams the records of its block range 32 at a time: blocks are only partly filled (survivors
        // of 32 candidates), so each step packs the unread records of up to 4 consecutive same-face blocks onto the
        // lanes -- (k, o) = current block and records of it already consumed.
        // block headers of the warp's range: ONE coalesced load per 32 blocks, kept in registers and read with shuffles
        // (a header load inside the loop would put a second global-memory latency in front of every record load)
        uint32_t k = kbeg, kwin = kbeg;
        uint32_t hw = (kbeg + lane < kend) ? __ldg(hdrs + kbeg + lane) : 0u;
        int o = 0;
        // plan(): the next step of the stream -- its face, whether this lane carries a record, and where the record is --
        // and advance (k, o).  Planning runs one step AHEAD of the arithmetic so that the step's three 16-byte lines per lane
        // can be prefetched into L1 while the previous step is being processed (the kernel was bound by the latency of
        // these loads: 50 % of its stall samples, profiles/r02_*bwd2*).
        auto plan = [&](int& f_out, bool& act_out, const float4*& src_out) -> bool {
            while (k < kend) {
                if (k >= kwin + 32u) {  // warp-uniform
                    kwin = k;
                    hw = (k + lane < kend) ? __ldg(hdrs + k + lane) : 0u;
                }
                const uint32_t wend = min(kend, kwin + 32u);  // chains stop at the header window
                const int rel = (int)(k - kwin);
                const uint32_t h0 = __shfl_sync(0xffffffffu, hw, rel), h1 = __shfl_sync(0xffffffffu, hw, (rel + 1) & 31);
                const uint32_t h2 = __shfl_sync(0xffffffffu, hw, (rel + 2) & 31), h3 = __shfl_sync(0xffffffffu, hw, (rel + 3) & 31);
                const int a0 = (int)(h0 >> 16) - o;
                if (a0 <= 0) { ++k; o = 0; continue; }  // block exhausted / empty (warp-uniform)
                const int f = (int)(h0 & 0xffffu);
                // records available in the following blocks while they belong to the same face (an empty block is transparent)
                int a1 = 0, a2 = 0, a3 = 0, nchain = 1;
                if (k + 1 < wend && ((h1 >> 16) == 0 || (int)(h1 & 0xffffu) == f)) {
                    a1 = (int)(h1 >> 16); nchain = 2;
                    if (k + 2 < wend && ((h2 >> 16) == 0 || (int)(h2 & 0xffffu) == f)) {
                        a2 = (int)(h2 >> 16); nchain = 3;
                        if (k + 3 < wend && ((h3 >> 16) == 0 || (int)(h3 & 0xffffu) == f)) { a3 = (int)(h3 >> 16); nchain = 4; }
                    }
                }
                const int p1 = a0, p2 = a0 + a1, p3 = p2 + a2, p4 = p3 + a3;
                int bi, pos;
                if (lane < p1) { bi = 0; pos = lane + o; }
                else if (lane < p2) { bi = 1; pos = lane - p1; }
                else if (lane < p3) { bi = 2; pos = lane - p2; }
                else { bi = 3; pos = lane - p3; }
                f_out = f;
                act_out = lane < p4;
                src_out = pb.recs + (size_t)(seg + 2 + k + bi) * BLK_F4 + pos;
                // advance the stream by min(32, p4) records: blocks consumed completely, then a partial one
                int left = min(32, p4), i = 0;
                const int av[4] = {a0, a1, a2, a3};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (i == q && q < nchain && left >= av[q]) { left -= av[q]; ++i; }
                }
                k += (uint32_t)i;
                o = (i == 0) ? o + left : ((i < nchain) ? left : 0);
                return true;
            }
            return false;
        };
        // The records of step i + 1 are prefetched into L1 while step i is processed, so the loads' latency is covered by a
        // whole step of arithmetic.  -DUMR_BWD2_REGPIPE=1 loads them into 12 registers instead: measured equal.
        auto load3 = [&](bool act, const float4* p, float4& a, float4& b_, float4& c) {
#if UMR_BWD2_REGPIPE
            if (act) { a = __ldg(p); b_ = __ldg(p + 32); c = __ldg(p + 64); }
#else
            if (act) {
                asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(p + 32));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(p + 64));
            }
#endif
        };
        int f_c = 0, f_n = 0;
        bool act_c = false, act_n = false;
        const float4 *src_c = nullptr, *src_n = nullptr;
        float4 c0r = make_float4(0, 0, 0, 0), c1r = c0r, c2r = c0r, n0r = c0r, n1r = c0r, n2r = c0r;
        bool have = plan(f_c, act_c, src_c);
        if (have) load3(act_c, src_c, c0r, c1r, c2r);
        while (have) {
            const bool have_n = plan(f_n, act_n, src_n);
            if (have_n) load3(act_n, src_n, n0r, n1r, n2r);
            const int f = f_c;
            if (f != cur_f) { flush(); cur_f = f; }
            int pre_tix = -1;   // PRE: this lane's texel of face f and its NC gradient terms (set below when it contributes)
            float pre_v[NC];
#pragma unroll
            for (int c = 0; c < NC; ++c) pre_v[c] = 0.f;
            if (act_c) {
#if UMR_BWD2_REGPIPE
                const float4 r0 = c0r, r1 = c1r, r2 = c2r;
#else
                const float4* src = src_c;
                const float4 r0 = __ldg(src), r1 = __ldg(src + 32), r2 = __ldg(src + 64);
#endif
                const float D = r0.x, sdx = r0.y, sdy = r0.z, zn = r0.w;  // zn: normalised depth exactly as the forward formed it
                const uint32_t meta = __float_as_uint(r1.w);
                const int pix = TS == 16 ? (int)(meta & 0xffu) : (int)(meta & 0x3ffu);
                const int tix = TS == 16 ? (int)((meta >> 8) & 0xffffu) : (int)((meta >> 10) & 0x3fffu);
                const bool front = (meta >> 24) & 1u;
                const float* sp = &s_pix[0][pix];
                float Cxy = 0.f;
                if (GEOM) {
                    const float g3 = sp[NC * NP];  // gradient of alpha
                    const float one_m_a = 1 - sp[(2 * NC + 1) * NP];
                    // g3 * ((1 - alpha) / max(1 - D, 1e-6)) (kernel.cu:584), in fp32 (the reference promotes to double;
                    // gradients are compared at 1e-4, see DESIGN.md "backward arithmetic")
                    Cxy = (one_m_a == 0.f || g3 == 0.f) ? g3 * one_m_a : g3 * __fdividef(one_m_a, fmaxf(1 - D, 1e-6f));
                }
                if (RGB == 0) {
                    if ((float)f == sp[(NV - 1) * NP]) {  // aggrs[1] = winning face id (:596)
                        if (TEXGRAD) {
                            float* gt = gtex_img + ((size_t)f * K.T2 + tix) * NC;
                            float g_[NC];
#pragma unroll
                            for (int c = 0; c < NC; ++c) g_[c] = sp[c * NP];
                            red_add_texel<NC>(gt, g_);
                        }
                    }
                } else if (front || K.double_side) {
                    float g[NC];
                    bool any = false;
#pragma unroll
                    for (int c = 0; c < NC; ++c) { g[c] = sp[c * NP]; any = any || g[c] != 0.f; }
                    if (any) {
                        const float s = __fdividef(D * expf((zn - sp[(NV - 1) * NP]) * K.r_gamma), sp[(NV - 2) * NP]);  // :608
                        if (s != 0.f) {
                            const size_t to = ((size_t)f * K.T2 + tix) * NC;
                            if (TEXGRAD) {
                                if (PRE) {
                                    pre_tix = tix;
#pragma unroll
                                    for (int c = 0; c < NC; ++c) pre_v[c] = s * g[c];
                                } else {
                                    float sg[NC];
#pragma unroll
                                    for (int c = 0; c < NC; ++c) sg[c] = s * g[c];
                                    red_add_texel<NC>(gtex_img + to, sg);
                                }
                            }
                            float Crgb = 0.f;
                            if (GEOM) {
#pragma unroll
                                for (int c = 0; c < NC; ++c) Crgb += g[c] * (__ldg(tex_img + to + c) - sp[(NPL + c) * NP]);
                                Crgb *= s;
                            }
                            if (GEOM && Crgb != 0.f) {
                                Cxy += __fdividef(Crgb, D);
                                const float zp = K.far_ - zn * (K.far_ - K.near_);
                                const float Cz = Crgb * K.r_gamma * K.r_nf * zp * zp;  // :624
                                acc[2] += Cz * r2.x;
                                acc[5] += Cz * r2.y;
                                acc[8] += Cz * r2.z;
                            }
                        }
                    }
                }
                if (GEOM) {
                    Cxy *= D * (1 - D) * K.r_sigma;  // :632
                    const float q = 2 * Cxy;          // :640 (the sign rides in sdx / sdy)
                    acc[0] += q * r1.x * sdx;
                    acc[1] += q * r1.x * sdy;
                    acc[3] += q * r1.y * sdx;
                    acc[4] += q * r1.y * sdy;
                    acc[6] += q * r1.z * sdx;
                    acc[7] += q * r1.z * sdy;
                }
            }
            if (PRE && TEXGRAD && RGB == 1) {  // all 32 lanes are here (the stream loop is warp-uniform)
                const bool has = pre_tix >= 0;
                if (__any_sync(0xffffffffu, has)) {
                    const uint32_t lt_ = (1u << lane) - 1u;
                    const uint32_t pm = __match_any_sync(0xffffffffu, has ? (uint32_t)pre_tix : (0x80000000u | (uint32_t)lane));
                    const bool leader = (pm & lt_) == 0u;
                    uint32_t rest = (has && leader) ? (pm & ~(1u << lane)) : 0u;  // the other lanes on this leader's texel
                    float sum[NC];
#pragma unroll
                    for (int c = 0; c < NC; ++c) sum[c] = pre_v[c];
                    while (__any_sync(0xffffffffu, rest != 0u)) {
                        const int src = rest ? __ffs(rest) - 1 : lane;
                        const bool take = rest != 0u;
                        rest &= rest - 1u;
#pragma unroll
                        for (int c = 0; c < NC; ++c) {
                            const float t = __shfl_sync(0xffffffffu, pre_v[c], src);
                            if (take) sum[c] += t;
                        }
                    }
                    if (has && leader) {
                        float* gt = gtex_img + ((size_t)f * K.T2 + pre_tix) * NC;
                        red_add_texel<NC>(gt, sum);
                    }
                }
            }
            f_c = f_n; act_c = act_n; src_c = src_n; have = have_n;
            c0r = n0r; c1r = n1r; c2r = n2r;
        }
        seg = next;
    }
    flush();
}

}  // namespace umr
