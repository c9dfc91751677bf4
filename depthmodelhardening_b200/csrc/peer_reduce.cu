// The one collective of stage 1 as ONE kernel over NVLink peer memory: the all-reduce of the shared
// patch gradient (with the scalar attack loss in its tail), the 1 / world averaging and -- optionally -- the L-inf
// PGD update (phy_obj_atk.py:98-100) that consumes it, without NCCL:
//
//   * every rank's patch_apply_bwd kernel accumulates its partial gradient into a buffer of a SYMMETRIC allocation
//     (same size on every rank, every rank's copy mapped into every process: CUDA VMM handles exchanged once by the
//     host -- torch.distributed._symmetric_memory is the plumbing);
//   * this kernel posts "my gradient of step e is complete" into every peer's flag block (system-scope release
//     store), waits for the same flag of every peer (acquire loads), then every CTA reads its slice of ALL ranks'
//     buffers straight over NVLink (volatile 128-bit loads: no L1 staleness) and adds them IN RANK ORDER -- every
//     rank computes the same bits, so the sign / Adam / threshold update that follows keeps the universal patch
//     bit-identical across ranks (what dist.allreduce_patch_grad guaranteed through NCCL);
//   * the last CTA of the rank posts "done reading" to the peers and waits for theirs before the kernel retires: when
//     the kernel is complete the rank's buffer is free for the next step's backward.  The step counter lives in
//     device memory (the kernel increments it), so the launch is CUDA-graph capturable and replayable.
//
// A step moves (world - 1) x 0.94 MB per rank over NVSwitch; the cost is the two flag exchanges (~2 x NVLink latency).
#include "../../include/dmh_b200.h"
#include "dmh_common.cuh"

using namespace dmh;

namespace {

#define PR_MAX_WORLD 16
#define PR_THREADS 256

struct PeerParams {
    const float* buf[PR_MAX_WORLD];      // every rank's gradient buffer as mapped in THIS process (buf[rank] = own)
    unsigned* flags[PR_MAX_WORLD];       // every rank's flag block: [0, world) ready[src], [world, 2 world) done[src]
    unsigned* state;                     // local: [0] step counter, [1] CTA arrival counter
    float* out;                          // local: the reduced, scaled gradient
    const float* adv;                    // optional fused L-inf update (all NULL: none)
    const float* clean;
    float* adv_out;
    long long n, n_update;
    float scale, alpha, eps;
    int rank, world;
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_volatile4(const float* p) {
    float4 v;
    asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float ld_volatile1(const float* p) {
    float v;
    asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float sgnf(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }

__device__ __forceinline__ float linf_update(float a, float g, float c, float alpha, float eps) {
    const float x = add_rn(a, mul_rn(alpha, sgnf(g)));
    const float delta = fminf(fmaxf(sub_rn(x, c), -eps), eps);
    return fminf(fmaxf(add_rn(c, delta), 0.0f), 1.0f);
}

__global__ void __launch_bounds__(PR_THREADS)
peer_allreduce_kernel(const PeerParams p) {
    const int tid = threadIdx.x;
    const int world = p.world;
    // step number of this launch: the counter is advanced by the last CTA at the very end, so every CTA reads the
    // same value here
    const unsigned e = *reinterpret_cast<volatile unsigned*>(p.state) + 1u;
    if (blockIdx.x == 0 && tid < world) {
        // the gradient was written by the previous kernel(s) of this stream: complete at this kernel's start;
        // the fence orders it before the flag for observers at system scope
        __threadfence_system();
        st_release_sys(p.flags[tid] + p.rank, e);
    }
    if (tid < world) {
        const unsigned* mine = p.flags[p.rank] + tid;
        while ((int)(ld_acquire_sys(mine) - e) < 0) { }
    }
    __syncthreads();

    const long long n4 = p.n >> 2;
    for (long long i = (long long)blockIdx.x * PR_THREADS + tid; i < n4; i += (long long)gridDim.x * PR_THREADS) {
        // the loads of (up to) 8 ranks are all in flight before the first addition; the sum is taken in rank order
        float4 v[8];
#pragma unroll
        for (int r = 0; r < 8; ++r)
            if (r < world) v[r] = ld_volatile4(p.buf[r] + 4 * i);
        float4 s = v[0];
#pragma unroll
        for (int r = 1; r < 8; ++r)
            if (r < world) { s.x = add_rn(s.x, v[r].x); s.y = add_rn(s.y, v[r].y); s.z = add_rn(s.z, v[r].z); s.w = add_rn(s.w, v[r].w); }
        for (int r = 8; r < world; ++r) {
            const float4 w = ld_volatile4(p.buf[r] + 4 * i);
            s.x = add_rn(s.x, w.x); s.y = add_rn(s.y, w.y); s.z = add_rn(s.z, w.z); s.w = add_rn(s.w, w.w);
        }
        s.x = mul_rn(s.x, p.scale); s.y = mul_rn(s.y, p.scale); s.z = mul_rn(s.z, p.scale); s.w = mul_rn(s.w, p.scale);
        reinterpret_cast<float4*>(p.out)[i] = s;
        if (p.adv_out && 4 * i + 3 < p.n_update) {
            const float4 a = reinterpret_cast<const float4*>(p.adv)[i], c = reinterpret_cast<const float4*>(p.clean)[i];
            float4 o;
            o.x = linf_update(a.x, s.x, c.x, p.alpha, p.eps); o.y = linf_update(a.y, s.y, c.y, p.alpha, p.eps);
            o.z = linf_update(a.z, s.z, c.z, p.alpha, p.eps); o.w = linf_update(a.w, s.w, c.w, p.alpha, p.eps);
            reinterpret_cast<float4*>(p.adv_out)[i] = o;
        }
    }
    // tail (n % 4 elements, and update elements whose float4 straddles n_update): one thread of the first CTA
    if (blockIdx.x == 0 && tid == 0) {
        for (long long i = 4 * n4; i < p.n; ++i) {
            float s = ld_volatile1(p.buf[0] + i);
            for (int r = 1; r < world; ++r) s = add_rn(s, ld_volatile1(p.buf[r] + i));
            p.out[i] = mul_rn(s, p.scale);
        }
    }
    if (p.adv_out) {
        const long long u0 = (p.n_update >> 2) << 2;       // update elements not covered by a whole float4 above
        __syncthreads();
        if (blockIdx.x == 0 && tid == 0) {
            __threadfence();
            for (long long i = u0; i < p.n_update; ++i) {
                float s = ld_volatile1(p.buf[0] + i);
                for (int r = 1; r < world; ++r) s = add_rn(s, ld_volatile1(p.buf[r] + i));
                p.adv_out[i] = linf_update(p.adv[i], mul_rn(s, p.scale), p.clean[i], p.alpha, p.eps);
            }
        }
    }

    // completion: the last CTA of this rank tells the peers that their buffers are no longer read from here, waits
    // for the same from them (then this rank's buffer is free for the next backward), and advances the step counter
    __shared__ int is_last;
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        is_last = atomicAdd(p.state + 1, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (is_last) {
        if (tid < world) {
            st_release_sys(p.flags[tid] + world + p.rank, e);
            const unsigned* mine = p.flags[p.rank] + world + tid;
            while ((int)(ld_acquire_sys(mine) - e) < 0) { }
        }
        __syncthreads();
        if (tid == 0) {
            p.state[1] = 0u;
            __threadfence();
            *reinterpret_cast<volatile unsigned*>(p.state) = e;
        }
    }
}

}  // namespace

extern "C" int dmh_peer_allreduce(const float* const* peer_bufs_host, unsigned* const* peer_flags_host, int rank, int world,
                                  long long n, float scale, float* out, unsigned* state, const float* adv,
                                  const float* clean, long long n_update, float alpha, float eps, float* adv_out,
                                  dmh_stream_t stream) {
    DMH_REQUIRE(peer_bufs_host && peer_flags_host && out && state, "dmh_peer_allreduce: null argument");
    DMH_REQUIRE(world >= 1 && world <= PR_MAX_WORLD && rank >= 0 && rank < world,
                "dmh_peer_allreduce: rank %d / world %d outside [1,%d]", rank, world, PR_MAX_WORLD);
    DMH_REQUIRE(n > 0, "dmh_peer_allreduce: n <= 0");
    DMH_REQUIRE(!adv_out || (adv && clean && n_update > 0 && n_update <= n),
                "dmh_peer_allreduce: the fused L-inf update needs adv, clean and 0 < n_update <= n");
    PeerParams p;
    memset(&p, 0, sizeof(p));
    for (int r = 0; r < world; ++r) {
        DMH_REQUIRE(peer_bufs_host[r] && peer_flags_host[r], "dmh_peer_allreduce: null buffer / flags of rank %d", r);
        DMH_REQUIRE((uintptr_t)peer_bufs_host[r] % 16 == 0, "dmh_peer_allreduce: buffer of rank %d not 16-byte aligned", r);
        p.buf[r] = peer_bufs_host[r]; p.flags[r] = peer_flags_host[r];
    }
    DMH_REQUIRE((uintptr_t)out % 16 == 0, "dmh_peer_allreduce: out not 16-byte aligned");
    DMH_REQUIRE(!adv_out || ((uintptr_t)adv % 16 == 0 && (uintptr_t)clean % 16 == 0 && (uintptr_t)adv_out % 16 == 0),
                "dmh_peer_allreduce: adv / clean / adv_out not 16-byte aligned");
    p.state = state; p.out = out; p.adv = adv; p.clean = clean; p.adv_out = adv_out;
    p.n = n; p.n_update = adv_out ? n_update : 0; p.scale = scale; p.alpha = alpha; p.eps = eps;
    p.rank = rank; p.world = world;
    // one float4 per thread where the buffer allows; every CTA spins on the peers' flags on its own (no CTA waits for
    // another CTA of this grid), so residency is not required for progress
    int grid = ceil_div(n >> 2, (long long)PR_THREADS);
    grid = grid < 1 ? 1 : (grid > 296 ? 296 : grid);
    DMH_LAUNCH(peer_allreduce_kernel, grid, PR_THREADS, 0, (cudaStream_t)stream)(p);
    DMH_CHECK_LAUNCH("dmh_peer_allreduce");
    return DMH_OK;
}
