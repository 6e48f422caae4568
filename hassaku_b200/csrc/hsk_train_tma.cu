// hsk_mf_train_fused, bulk-copy (TMA, UBLKCP) ring kernel for bpr / bce: CTA = one sample, 4 warps, every warp owns a
// ring of item-row buffers in shared memory filled by 1-D cp.async.bulk with per-stage mbarriers; the user row and the
// positive row are bulk-copied once per CTA; dot by warp shuffle, loss + dL/ds in registers, dL/du in registers, the item
// row gradients leave as 128-bit vector reductions (RED.E.ADD.F32x4).
// This is the second cut of the inner loop (round 2 A/B on a B200, cfg2 shape: 151.6 us vs 155.6 us for the first cut,
// which was deleted; 174 instead of 249 SASS instructions per item row):
//   * shared-memory base addresses, flags and 1 / count are pinned in registers (`asm volatile` moves): the compiler
//     otherwise rebuilds the shared::cluster window address (S2R SR_CgaCtaId + LEA + IMAD) at every use, ~30 instr / row;
//   * lane t keeps the BYTE OFFSET of its slot's row (idx * ld * 4) next to the index: the row address for the bulk copy
//     and for the REDs is one 64-bit add of a shuffled offset instead of a 64-bit multiply each;
//   * the per-row lane-0 work (item-bias RED, scores_out / dscores_out stores, their address arithmetic and branches)
//     leaves the loop: every lane latches its own slot's score and dL/ds when the warp reaches it, and after the loop all
//     lanes issue their bias RED / stores at once (coalesced);
//   * 1 / (1 + e) and log(1 + e) with the argument in (1, 2]: rcp.approx / lg2.approx directly (1 ulp; no range checks,
//     no slow path);
//   * the item row is read from shared memory ONCE (dot product and dL/du accumulation share the registers).
#include "hsk_train.cuh"

namespace hsk {

namespace {

__device__ __forceinline__ uint32_t pin_u32(uint32_t x) {
    uint32_t y;
    asm volatile("mov.u32 %0, %1;" : "=r"(y) : "r"(x));
    return y;
}
__device__ __forceinline__ float pin_f32(float x) {
    float y;
    asm volatile("mov.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {   // x in [1, 2]: 1 ulp
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_approx(float x) {   // x in [1, 2]
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float4 lds128_nv(uint32_t addr) {   // not volatile: the loaded row is reused from registers
    float4 v;
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ float lds32_nv(uint32_t addr) {
    float v;
    asm("ld.shared.f32 %0, [%1];\n" : "=f"(v) : "r"(addr) : "memory");
    return v;
}

enum : uint32_t { F_UB = 1, F_IB = 2, F_GB = 4, F_GIB = 8, F_SC = 16, F_DS = 32, F_RED = 64 };

template <int NV, bool TAILS>
__device__ __forceinline__ void row_load(Row<NV>& r, uint32_t base /* includes lane * 16 */, uint32_t lane) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        if (TAILS && k == NV - 1) r.v[k] = make_float4(lds32_nv(base - lane * 12 + 512 * k), 0.f, 0.f, 0.f);
        else r.v[k] = lds128_nv(base + 512 * k);
    }
}
template <int NV, bool TAILS>
__device__ __forceinline__ float row_dot(const Row<NV>& u, const Row<NV>& v) {
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        a0 = fmaf(u.v[k].x, v.v[k].x, a0);
        if (!(TAILS && k == NV - 1)) {
            a1 = fmaf(u.v[k].y, v.v[k].y, a1);
            a0 = fmaf(u.v[k].z, v.v[k].z, a0);
            a1 = fmaf(u.v[k].w, v.v[k].w, a1);
        }
    }
    return a0 + a1;
}
template <int NV, bool TAILS>
__device__ __forceinline__ void row_axpy(Row<NV>& g, float a, const Row<NV>& v) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        g.v[k].x = fmaf(a, v.v[k].x, g.v[k].x);
        if (!(TAILS && k == NV - 1)) {
            g.v[k].y = fmaf(a, v.v[k].y, g.v[k].y);
            g.v[k].z = fmaf(a, v.v[k].z, g.v[k].z);
            g.v[k].w = fmaf(a, v.v[k].w, g.v[k].w);
        }
    }
}
// dst_row (byte pointer to the row start) += a * u
template <int NV, bool TAILS>
__device__ __forceinline__ void row_red(const Row<NV>& u, char* dst_row, float a, int nvec, uint32_t lane) {
    float4* p = reinterpret_cast<float4*>(dst_row) + lane;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        if (TAILS && k == NV - 1) {
            if (128 * k + (int)lane < nvec * 4) atomicAdd(reinterpret_cast<float*>(dst_row) + 128 * k + lane, a * u.v[k].x);
        } else if (k < NV - 1 || (int)lane + 32 * k < nvec) {
            atomicAdd(p + 32 * k, make_float4(a * u.v[k].x, a * u.v[k].y, a * u.v[k].z, a * u.v[k].w));
        }
    }
}

constexpr int min_blocks_ring(int nv) { return nv <= 4 ? 6 : (nv <= 6 ? 5 : 4); }

}  // namespace

template <int NV, int LOSS, int STAGES, bool TAILS>
__global__ void __launch_bounds__(kWarpsPerCta * 32, min_blocks_ring(NV)) mf_train_fused_tma_kernel(TrainArgs a) {
    extern __shared__ __align__(128) unsigned char dyn[];
    __shared__ uint64_t bar_u, bar_v0;
    __shared__ uint64_t bars[kWarpsPerCta][STAGES];
    __shared__ float sm_ds0[kWarpsPerCta], sm_dsum[kWarpsPerCta], sm_loss[kWarpsPerCta];
    constexpr int kSlot = NV * 512;

    const int warp = threadIdx.x >> 5;
    const uint32_t lane = threadIdx.x & 31;
    const int b = blockIdx.x;
    const int ld = a.ld, nvec = a.nvec, N1 = a.N1;
    const int64_t u = a.u_idx[b];
    const int64_t* __restrict__ irow = a.i_idx + (int64_t)b * N1;
    const int64_t i0 = irow[0];
    if (bad_index(u, a.n_users) || (LOSS == HSK_LOSS_BPR && bad_index(i0, a.n_items))) {
        if (threadIdx.x == 0 && a.status) atomicOr(a.status, HSK_STATUS_BAD_INDEX);
        return;
    }
    const uint32_t row_bytes = pin_u32((uint32_t)ld * 4u);
    unsigned char* slot_u = dyn;
    unsigned char* slot_v0 = dyn + kSlot;
    unsigned char* ring = dyn + kSlot * (2 + warp * STAGES);
    {
        const int tail_f4 = NV * 32 - nvec;
        const int n_slots = 2 + kWarpsPerCta * STAGES;
        for (int e = threadIdx.x; e < n_slots * tail_f4; e += kWarpsPerCta * 32) {
            const int sl = e / tail_f4, o = e - sl * tail_f4;
            reinterpret_cast<float4*>(dyn + sl * kSlot)[nvec + o] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    if (threadIdx.x == 0) { mbar_init(&bar_u, 1); mbar_init(&bar_v0, 1); }
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&bars[warp][s], 1);
    }
    mbar_fence_init();
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar_u, row_bytes);
        bulk_g2s(slot_u, a.Uw + u * ld, row_bytes, &bar_u);
        if (LOSS == HSK_LOSS_BPR) {
            mbar_expect_tx(&bar_v0, row_bytes);
            bulk_g2s(slot_v0, a.Vw + i0 * ld, row_bytes, &bar_v0);
        }
    }

    // ---- this warp's item slots: j = jbase + warp + 4 t, t < n_my <= 32; lane t owns slot t ----
    const int first = (LOSS == HSK_LOSS_BPR) ? 1 : 0;
    const int jbase = first + blockIdx.y * a.j_per_cta;
    const int jend = min(N1, jbase + a.j_per_cta);
    const int n_my = max(0, (jend - jbase - warp + kWarpsPerCta - 1) / kWarpsPerCta);
    int64_t my_idx = 0;
    bool ok = false;
    if ((int)lane < n_my) {
        my_idx = irow[jbase + warp + kWarpsPerCta * (int)lane];
        ok = !bad_index(my_idx, a.n_items);
        if (!ok && a.status) atomicOr(a.status, HSK_STATUS_BAD_INDEX);
    }
    const int64_t my_off_v = ok ? my_idx * (int64_t)row_bytes : 0;   // byte offset of my slot's row in Vw / gV
    // pinned (two 32-bit halves): otherwise the 64-bit multiply is redone at every shuffle site
    const int64_t my_off = (int64_t)(((uint64_t)pin_u32((uint32_t)((uint64_t)my_off_v >> 32)) << 32) |
                                     (uint64_t)pin_u32((uint32_t)my_off_v));
    const float my_ib = (ok && a.Ib) ? __ldg(a.Ib + my_idx) : 0.f;
    const uint32_t valid = __ballot_sync(kFull, ok);
    const char* __restrict__ Vb = reinterpret_cast<const char*>(a.Vw);
    char* __restrict__ gVb = reinterpret_cast<char*>(a.gV);

    const uint32_t flags = pin_u32((a.Ub ? F_UB : 0u) | (a.Ib ? F_IB : 0u) | (a.Gb ? F_GB : 0u) | (a.gIb ? F_GIB : 0u) |
                                   (a.scores_out ? F_SC : 0u) | (a.dscores_out ? F_DS : 0u) | F_RED);
    const uint32_t ring_a = pin_u32(smem_u32(ring)), bars_a = pin_u32(smem_u32(&bars[warp][0]));
    const uint32_t bars_end = pin_u32(bars_a + 8 * STAGES);
    const uint32_t lane16 = lane * 16u;
    uint32_t to_issue = valid, to_consume = valid;
    uint32_t iss_slot = ring_a, iss_bar = bars_a;
    auto issue_next = [&]() {
        const int t = __ffs(to_issue) - 1;
        to_issue &= to_issue - 1;
        const int64_t off = __shfl_sync(kFull, my_off, t);
        if (lane == 0) {
            mbar_expect_tx_a(iss_bar, row_bytes);
            bulk_g2s_a(iss_slot, Vb + off, row_bytes, iss_bar);
        }
        iss_slot += kSlot; iss_bar += 8;
        if (iss_bar == bars_end) { iss_slot = ring_a; iss_bar = bars_a; }
    };
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s)
        if (to_issue) issue_next();

    Row<NV> ur, gu;
    gu.zero();
    mbar_wait(&bar_u, 0);
    row_load<NV, TAILS>(ur, smem_u32(slot_u) + lane16, lane);

    const float ubv = (flags & F_UB) ? a.Ub[u] : 0.f, gbv = (flags & F_GB) ? a.Gb[0] : 0.f;
    const float bias_ug = pin_f32(ubv + gbv);       // only used when neither or one of them is present (see below)
    float s0 = 0.f;
    if (LOSS == HSK_LOSS_BPR) {
        mbar_wait(&bar_v0, 0);
        Row<NV> v0;
        row_load<NV, TAILS>(v0, smem_u32(slot_v0) + lane16, lane);
        s0 = warp_sum(row_dot<NV, TAILS>(ur, v0));
        if (flags & F_UB) s0 += ubv;     // sgd_alg.py:173-178 order
        if (flags & F_IB) s0 += __ldg(a.Ib + i0);
        if (flags & F_GB) s0 += gbv;
    }
    (void)bias_ug;

    float ds0 = 0.f, dsum = 0.f, loss_local = 0.f;
    float my_sc = 0.f, my_ds = 0.f;                  // score and dL/ds of MY slot, latched when the warp reaches it
    const float invf = pin_f32((float)a.inv_count);
    uint32_t cur_slot = ring_a + lane16, cur_bar = bars_a;
    uint32_t parity = 0;
    while (to_consume) {
        if (to_issue) issue_next();
        const int t = __ffs(to_consume) - 1;
        to_consume &= to_consume - 1;
        const int64_t off = __shfl_sync(kFull, my_off, t);
        const float ib = __shfl_sync(kFull, my_ib, t);
        mbar_wait_a(cur_bar, parity);
        Row<NV> vr;
        row_load<NV, TAILS>(vr, cur_slot, lane);
        float sj = warp_sum(row_dot<NV, TAILS>(ur, vr));
        if (flags & F_UB) sj += ubv;     // sgd_alg.py:173-178 order
        if (flags & F_IB) sj += ib;
        if (flags & F_GB) sj += gbv;
        float dsj;
        if (LOSS == HSK_LOSS_BPR) {
            // sigma(x) - 1 = -1 / (1 + e^x); one exp serves the gradient and the loss (see hsk_train_tma.cu)
            const float x = s0 - sj;
            const float e = expf(-fabsf(x));
            const float r = rcp_approx(1.f + e);
            const float dx = -(x >= 0.f ? e * r : r) * invf;
            dsj = -dx;
            ds0 += dx;
            loss_local += (lg2_approx(1.f + e) * 0.6931471805599453f - fminf(x, 0.f)) * invf;
        } else {
            const float y = (warp + kWarpsPerCta * t + jbase == 0) ? 1.f : 0.f;
            const float e = expf(-fabsf(sj));
            const float r = rcp_approx(1.f + e);
            const float sig = sj >= 0.f ? r : e * r;
            dsj = (sig - y) * invf;
            loss_local += ((1.f - y) * sj + lg2_approx(1.f + e) * 0.6931471805599453f - fminf(sj, 0.f)) * invf;
        }
        dsum += dsj;
        if ((int)lane == t) { my_sc = sj; my_ds = dsj; }
        row_axpy<NV, TAILS>(gu, dsj, vr);
        if (flags & F_RED) row_red<NV, TAILS>(ur, gVb + off, dsj, nvec, lane);
        cur_slot += kSlot; cur_bar += 8;
        if (cur_bar == bars_end) { cur_slot = ring_a + lane16; cur_bar = bars_a; parity ^= 1u; }
        __syncwarp();
    }
    // ---- per-slot scalars, all lanes at once: item-bias gradient, scores_out, dscores_out ----
    if (ok) {
        if (flags & F_GIB) atomicAdd(a.gIb + my_idx, my_ds);
        if (flags & (F_SC | F_DS)) {
            const int64_t o = (int64_t)b * N1 + jbase + warp + kWarpsPerCta * (int)lane;
            if (flags & F_SC) a.scores_out[o] = my_sc;
            if (flags & F_DS) a.dscores_out[o] = my_ds;
        }
    }

    // ---- combine the four warps (their rings are idle now: reuse them as the reduction buffer) ----
    float4* red = reinterpret_cast<float4*>(ring);
#pragma unroll
    for (int k = 0; k < NV; ++k) red[k * 32 + lane] = gu.v[k];
    if (lane == 0) { sm_ds0[warp] = ds0; sm_dsum[warp] = dsum; sm_loss[warp] = loss_local; }
    __syncthreads();
    if (warp == 0) {
        float d0 = 0.f, dsm = 0.f;
        double ls = 0.0;
#pragma unroll
        for (int w = 0; w < kWarpsPerCta; ++w) { d0 += sm_ds0[w]; dsm += sm_dsum[w]; ls += (double)sm_loss[w]; }
#pragma unroll
        for (int w = 1; w < kWarpsPerCta; ++w) {
            const float4* o = reinterpret_cast<const float4*>(dyn + kSlot * (2 + w * STAGES));
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                const float4 x = o[k * 32 + lane];
                gu.v[k].x += x.x; gu.v[k].y += x.y; gu.v[k].z += x.z; gu.v[k].w += x.w;
            }
        }
        const int64_t rowoff = (int64_t)b * N1;
        if (LOSS == HSK_LOSS_BPR) {
            Row<NV> v0;
            row_load<NV, TAILS>(v0, smem_u32(slot_v0) + lane16, lane);
            row_axpy<NV, TAILS>(gu, d0, v0);
            if (flags & F_RED) row_red<NV, TAILS>(ur, gVb + i0 * (int64_t)row_bytes, d0, nvec, lane);
        }
        row_red<NV, TAILS>(gu, reinterpret_cast<char*>(a.gU) + u * (int64_t)row_bytes, 1.0f, nvec, lane);
        if (lane == 0) {
            if (LOSS == HSK_LOSS_BPR) {
                if (a.gIb) atomicAdd(a.gIb + i0, d0);
                dsm += d0;
                if (blockIdx.y == 0 && a.scores_out) a.scores_out[rowoff] = s0;
                if (a.dscores_out) { if (gridDim.y == 1) a.dscores_out[rowoff] = d0; else atomicAdd(a.dscores_out + rowoff, d0); }
            }
            if (a.gUb) atomicAdd(a.gUb + u, dsm);
            if (a.gGb) atomicAdd(a.gGb, dsm);
            if (a.loss_accum && ls != 0.0) atomicAdd(a.loss_accum, ls);
        }
    }
}

namespace {

template <int NV, int LOSS, int STAGES>
int launch_one2(const TrainArgs& a, dim3 grid, cudaStream_t s) {
    const size_t smem = (size_t)NV * 512 * (2 + kWarpsPerCta * STAGES);
    const bool tails = (a.nvec - 32 * (NV - 1)) <= 8;
    auto kern = tails ? mf_train_fused_tma_kernel<NV, LOSS, STAGES, true> : mf_train_fused_tma_kernel<NV, LOSS, STAGES, false>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_err(HSK_ERR_CUDA, "hsk_mf_train_fused: smem attribute: %s", cudaGetErrorString(e));
    }
    kern<<<grid, kWarpsPerCta * 32, smem, s>>>(a);
    return check_launch("hsk_mf_train_fused(tma)");
}

template <int NV, int LOSS>
int launch_nv2(const TrainArgs& a, dim3 grid, cudaStream_t s) {
    if (NV <= 2) return launch_one2<NV, LOSS, 4>(a, grid, s);
    if (NV <= 4) return launch_one2<NV, LOSS, 3>(a, grid, s);
    return launch_one2<NV, LOSS, 2>(a, grid, s);
}

}  // namespace

int launch_train_fused_tma(const TrainArgs& a, int loss_kind, cudaStream_t s) {
    const int first = (loss_kind == HSK_LOSS_BPR) ? 1 : 0;
    const int n_slots = a.N1 - first;
    dim3 grid(a.B, (n_slots + a.j_per_cta - 1) / a.j_per_cta);
    const int nv = (a.nvec + 31) / 32;
    if (loss_kind == HSK_LOSS_BPR) {
        HSK_DISPATCH_NV(nv, return (launch_nv2<NV, HSK_LOSS_BPR>(a, grid, s)));
    } else {
        HSK_DISPATCH_NV(nv, return (launch_nv2<NV, HSK_LOSS_BCE>(a, grid, s)));
    }
    return HSK_OK;
}

}  // namespace hsk
