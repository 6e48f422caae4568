// hsk_mf_train_fused, pipelined variant (bpr / bce): gathered rows are staged in shared memory by the TMA engine.
//
// Why: the register-gather kernel (hsk_train.cu) keeps ONE row in flight per warp at 20 warps/SM and is latency
// bound (ncu r01a: 0.78 eligible warps/scheduler, L2 47 %, issue 50 %).  Here every warp owns a ring of STAGES row
// buffers in shared memory; lane 0 issues one `cp.async.bulk` (UBLKCP, 1-D bulk copy, 16 B aligned rows of ld * 4 bytes)
// per row, completion is signalled on a per-stage mbarrier (expect_tx), and the warp consumes row t while rows
// t+1 .. t+STAGES-1 are in flight — no registers are tied up by loads, so 32+ warps/SM stay resident.
// The user row and the positive item row are bulk-copied once per CTA into shared slots all four warps read.
// Index rows and item biases of a warp's slots are fetched up front with one coalesced load per lane.
#include "hsk_train.cuh"

namespace hsk {

// Shared-memory row slots are NV * 512 bytes: the TMA writes ld * 4 bytes, the tail [ld * 4, NV * 512) is zeroed once
// per CTA and never written again, so the dot / axpy loops need no bounds checks (the zero tail contributes nothing).
// TAILS: the last of the NV rounds is a SCALAR round (one float per lane) instead of a masked float4 round — used when
// at most 32 floats remain after the full rounds (d = 402: 3 full rounds + 20 floats), which saves the 3/4-empty fourth
// float4 round in every dot / axpy / reduction.
template <int NV, bool TAILS>
__device__ __forceinline__ void row_from_smem(Row<NV>& r, const float4* p, int lane) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        if (TAILS && k == NV - 1) r.v[k] = make_float4(reinterpret_cast<const float*>(p)[128 * k + lane], 0.f, 0.f, 0.f);
        else r.v[k] = p[lane + 32 * k];
    }
}
template <int NV, bool TAILS>
__device__ __forceinline__ float dot_smem(const Row<NV>& u, const float4* p, int lane) {
    float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        if (TAILS && k == NV - 1) {
            acc0 = fmaf(u.v[k].x, reinterpret_cast<const float*>(p)[128 * k + lane], acc0);
        } else {
            const float4 v = p[lane + 32 * k];
            acc0 = fmaf(u.v[k].x, v.x, acc0);
            acc1 = fmaf(u.v[k].y, v.y, acc1);
            acc0 = fmaf(u.v[k].z, v.z, acc0);
            acc1 = fmaf(u.v[k].w, v.w, acc1);
        }
    }
    return acc0 + acc1;
}
template <int NV, bool TAILS>
__device__ __forceinline__ void axpy_smem(Row<NV>& g, float a, const float4* p, int lane) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        if (TAILS && k == NV - 1) {
            g.v[k].x = fmaf(a, reinterpret_cast<const float*>(p)[128 * k + lane], g.v[k].x);
        } else {
            const float4 v = p[lane + 32 * k];
            g.v[k].x = fmaf(a, v.x, g.v[k].x);
            g.v[k].y = fmaf(a, v.y, g.v[k].y);
            g.v[k].z = fmaf(a, v.z, g.v[k].z);
            g.v[k].w = fmaf(a, v.w, g.v[k].w);
        }
    }
}
// dst += a * u: full rounds unguarded, only the last round is bounds-checked
template <int NV, bool TAILS>
__device__ __forceinline__ void red_row(const Row<NV>& u, float* __restrict__ dst, float a, int nvec, int lane) {
    float4* p = reinterpret_cast<float4*>(dst) + lane;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        if (TAILS && k == NV - 1) {
            if (128 * k + lane < nvec * 4) atomicAdd(dst + 128 * k + lane, a * u.v[k].x);
        } else if (k < NV - 1 || lane + 32 * k < nvec) {
            atomicAdd(p + 32 * k, make_float4(a * u.v[k].x, a * u.v[k].y, a * u.v[k].z, a * u.v[k].w));
        }
    }
}

template <int NV, bool TAILS>
__device__ __forceinline__ float dot_smem_a(const Row<NV>& u, uint32_t base) {   // base already includes lane * 16
    float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        if (TAILS && k == NV - 1) {
            acc0 = fmaf(u.v[k].x, lds32(base - (threadIdx.x & 31) * 12 + 512 * k), acc0);
        } else {
            const float4 v = lds128(base + 512 * k);
            acc0 = fmaf(u.v[k].x, v.x, acc0);
            acc1 = fmaf(u.v[k].y, v.y, acc1);
            acc0 = fmaf(u.v[k].z, v.z, acc0);
            acc1 = fmaf(u.v[k].w, v.w, acc1);
        }
    }
    return acc0 + acc1;
}
template <int NV, bool TAILS>
__device__ __forceinline__ void axpy_smem_a(Row<NV>& g, float a, uint32_t base) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        if (TAILS && k == NV - 1) {
            g.v[k].x = fmaf(a, lds32(base - (threadIdx.x & 31) * 12 + 512 * k), g.v[k].x);
        } else {
            const float4 v = lds128(base + 512 * k);
            g.v[k].x = fmaf(a, v.x, g.v[k].x);
            g.v[k].y = fmaf(a, v.y, g.v[k].y);
            g.v[k].z = fmaf(a, v.z, g.v[k].z);
            g.v[k].w = fmaf(a, v.w, g.v[k].w);
        }
    }
}

constexpr int min_blocks_for(int nv) { return nv <= 4 ? 7 : (nv <= 6 ? 5 : 4); }

template <int NV, int LOSS, int STAGES, bool TAILS>
__global__ void __launch_bounds__(kWarpsPerCta * 32, min_blocks_for(NV)) mf_train_fused_tma_kernel(TrainArgs a) {
    extern __shared__ __align__(128) unsigned char dyn[];
    __shared__ uint64_t bar_u, bar_v0;
    __shared__ uint64_t bars[kWarpsPerCta][STAGES];
    __shared__ float sm_ds0[kWarpsPerCta], sm_dsum[kWarpsPerCta], sm_loss[kWarpsPerCta];
    constexpr int kSlot = NV * 512;  // bytes per row slot

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x;
    const int ld = a.ld, nvec = a.nvec, N1 = a.N1;
    const int64_t u = a.u_idx[b];
    const int64_t* __restrict__ irow = a.i_idx + (int64_t)b * N1;
    const int64_t i0 = irow[0];
    if (bad_index(u, a.n_users) || (LOSS == HSK_LOSS_BPR && bad_index(i0, a.n_items))) {
        if (threadIdx.x == 0 && a.status) atomicOr(a.status, HSK_STATUS_BAD_INDEX);
        return;
    }
    const uint32_t row_bytes = (uint32_t)ld * 4u;
    unsigned char* slot_u = dyn;
    unsigned char* slot_v0 = dyn + kSlot;
    unsigned char* ring = dyn + kSlot * (2 + warp * STAGES);

    // zero the slot tails (one pass over the CTA's slots), init barriers
    {
        const int tail_f4 = NV * 32 - nvec;  // float4 per slot beyond the row
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

    // ---- this warp's item slots: j = jbase + warp + 4 t, t < n_my <= 32 ----
    const int first = (LOSS == HSK_LOSS_BPR) ? 1 : 0;
    const int jbase = first + blockIdx.y * a.j_per_cta;
    const int jend = min(N1, jbase + a.j_per_cta);
    const int n_my = max(0, (jend - jbase - warp + kWarpsPerCta - 1) / kWarpsPerCta);
    int64_t my_idx = 0;
    bool ok = false;
    if (lane < n_my) {
        my_idx = irow[jbase + warp + kWarpsPerCta * lane];
        ok = !bad_index(my_idx, a.n_items);
        if (!ok && a.status) atomicOr(a.status, HSK_STATUS_BAD_INDEX);
    }
    const float my_ib = (ok && a.Ib) ? __ldg(a.Ib + my_idx) : 0.f;
    const uint32_t valid = __ballot_sync(kFull, ok);
    const float* __restrict__ Vw = a.Vw;
    float* __restrict__ gV = a.gV;

    // shared-memory addresses of this warp's ring and barriers, advanced incrementally (no multiplies in the loop)
    const uint32_t ring_a = smem_u32(ring), bars_a = smem_u32(&bars[warp][0]);
    const uint32_t lane16 = (uint32_t)lane * 16u;
    uint32_t to_issue = valid, to_consume = valid;
    uint32_t iss_slot = ring_a, iss_bar = bars_a;
    auto issue_next = [&]() {
        const int t = __ffs(to_issue) - 1;
        to_issue &= to_issue - 1;
        const int64_t it = __shfl_sync(kFull, my_idx, t);
        if (lane == 0) {
            mbar_expect_tx_a(iss_bar, row_bytes);
            bulk_g2s_a(iss_slot, Vw + it * ld, row_bytes, iss_bar);
        }
        iss_slot += kSlot; iss_bar += 8;
        if (iss_bar == bars_a + 8 * STAGES) { iss_slot = ring_a; iss_bar = bars_a; }
    };
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s)
        if (to_issue) issue_next();

    Row<NV> ur, gu;
    gu.zero();
    mbar_wait(&bar_u, 0);
    row_from_smem<NV, TAILS>(ur, reinterpret_cast<const float4*>(slot_u), lane);

    float s0 = 0.f;
    if (LOSS == HSK_LOSS_BPR) {
        mbar_wait(&bar_v0, 0);
        s0 = warp_sum(dot_smem<NV, TAILS>(ur, reinterpret_cast<const float4*>(slot_v0), lane));
        if (a.Ub) s0 += a.Ub[u];
        if (a.Ib) s0 += __ldg(a.Ib + i0);
        if (a.Gb) s0 += a.Gb[0];
    }
    const bool has_ub = a.Ub != nullptr, has_ib = a.Ib != nullptr, has_gb = a.Gb != nullptr;
    const float ubv = has_ub ? a.Ub[u] : 0.f, gbv = has_gb ? a.Gb[0] : 0.f;

    float ds0 = 0.f, dsum = 0.f, loss_local = 0.f;
    const float invf = (float)a.inv_count;
    const bool write_scores = a.scores_out != nullptr, write_ds = a.dscores_out != nullptr;
    const bool do_red = !(a.debug_flags & 1);
    float* __restrict__ gIb = a.gIb;
    const int64_t rowoff = (int64_t)b * N1;
    uint32_t cur_slot = ring_a + lane16, cur_bar = bars_a;
    uint32_t parity = 0;
    while (to_consume) {
        if (to_issue) issue_next();
        const int t = __ffs(to_consume) - 1;
        to_consume &= to_consume - 1;
        const int64_t it = __shfl_sync(kFull, my_idx, t);
        const float ib = __shfl_sync(kFull, my_ib, t);
        mbar_wait_a(cur_bar, parity);
        float sj = warp_sum(dot_smem_a<NV, TAILS>(ur, cur_slot));
        if (has_ub) sj += ubv;     // sgd_alg.py:173-178 order
        if (has_ib) sj += ib;
        if (has_gb) sj += gbv;
        float dsj;
        if (LOSS == HSK_LOSS_BPR) {
            const float x = s0 - sj;
            // sigma(x) - 1 = -1 / (1 + e^x); one exp serves the gradient and the loss:
            //   x >= 0: e = e^-x, sig-1 = -e/(1+e), -logsig = log(1+e)
            //   x <  0: e = e^x,  sig-1 = -1/(1+e), -logsig = log(1+e) - x
            const float e = expf(-fabsf(x));
            const float r = __frcp_rn(1.f + e);
            const float dx = -(x >= 0.f ? e * r : r) * invf;  // dL/dx = (sigma(x) - 1) / (B N)
            dsj = -dx;
            ds0 += dx;
            // the reported loss only (not the gradient): log(1+e) by the fast log, absolute error ~1e-7 per term
            loss_local += (__logf(1.f + e) - fminf(x, 0.f)) * invf;
        } else {
            const float y = (warp + kWarpsPerCta * t + jbase == 0) ? 1.f : 0.f;
            const float e = expf(-fabsf(sj));
            const float r = __frcp_rn(1.f + e);
            const float sig = sj >= 0.f ? r : e * r;
            dsj = (sig - y) * invf;
            loss_local += ((1.f - y) * sj + __logf(1.f + e) - fminf(sj, 0.f)) * invf;
        }
        dsum += dsj;
        axpy_smem_a<NV, TAILS>(gu, dsj, cur_slot);
        if (do_red) red_row<NV, TAILS>(ur, gV + it * ld, dsj, nvec, lane);
        if (lane == 0) {
            if (gIb) atomicAdd(gIb + it, dsj);
            if (write_scores | write_ds) {
                const int j = jbase + warp + kWarpsPerCta * t;
                if (write_scores) a.scores_out[rowoff + j] = sj;
                if (write_ds) a.dscores_out[rowoff + j] = dsj;
            }
        }
        cur_slot += kSlot; cur_bar += 8;
        if (cur_bar == bars_a + 8 * STAGES) { cur_slot = ring_a + lane16; cur_bar = bars_a; parity ^= 1u; }
        __syncwarp();
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
        if (LOSS == HSK_LOSS_BPR) {
            axpy_smem<NV, TAILS>(gu, d0, reinterpret_cast<const float4*>(slot_v0), lane);
            if (do_red) red_row<NV, TAILS>(ur, gV + i0 * ld, d0, nvec, lane);
        }
        red_row<NV, TAILS>(gu, a.gU + u * ld, 1.0f, nvec, lane);
        if (lane == 0) {
            if (LOSS == HSK_LOSS_BPR) {
                if (a.gIb) atomicAdd(a.gIb + i0, d0);
                dsm += d0;
                if (blockIdx.y == 0 && write_scores) a.scores_out[rowoff] = s0;
                if (write_ds) { if (gridDim.y == 1) a.dscores_out[rowoff] = d0; else atomicAdd(a.dscores_out + rowoff, d0); }
            }
            if (a.gUb) atomicAdd(a.gUb + u, dsm);
            if (a.gGb) atomicAdd(a.gGb, dsm);
            if (a.loss_accum && ls != 0.0) atomicAdd(a.loss_accum, ls);
        }
    }
}

template <int NV, int LOSS, int STAGES>
static int launch_one(const TrainArgs& a, dim3 grid, int row_pad, cudaStream_t s) {
    (void)row_pad;
    const size_t smem = (size_t)NV * 512 * (2 + kWarpsPerCta * STAGES);
    // scalar tail round when at most 32 floats (8 float4) remain after NV - 1 full rounds
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
static int launch_nv(const TrainArgs& a, dim3 grid, int row_pad, cudaStream_t s) {
    // ring depth by row size: keep >= ~6 CTAs/SM resident within 227 KB of shared memory
    if (NV <= 2) return launch_one<NV, LOSS, 4>(a, grid, row_pad, s);
    if (NV <= 4) return launch_one<NV, LOSS, 3>(a, grid, row_pad, s);
    return launch_one<NV, LOSS, 2>(a, grid, row_pad, s);
}

int launch_train_fused_tma(const TrainArgs& a, int loss_kind, cudaStream_t s) {
    const int first = (loss_kind == HSK_LOSS_BPR) ? 1 : 0;
    const int n_slots = a.N1 - first;
    dim3 grid(a.B, (n_slots + a.j_per_cta - 1) / a.j_per_cta);
    const int row_pad = ((a.ld * 4 + 127) / 128) * 128;
    // every warp needs NV * 32 float4 of ring space for the final reduction
    const int nv = (a.nvec + 31) / 32;
    if (loss_kind == HSK_LOSS_BPR) {
        HSK_DISPATCH_NV(nv, return (launch_nv<NV, HSK_LOSS_BPR>(a, grid, row_pad, s)));
    } else {
        HSK_DISPATCH_NV(nv, return (launch_nv<NV, HSK_LOSS_BCE>(a, grid, row_pad, s)));
    }
    return HSK_OK;
}

}  // namespace hsk
