/*
 * gg_kernels_v4.cuh -- the round-1 production kernel (one shared-memory stage per CTA, CTA-wide staging of the
 * result rows), kept as CFDP_KERNEL=3 for side-by-side measurements against the round-2 kernel in gg_kernels.cuh.
 *
 * Replaces private_compute_gradients_gg (reference src/gradients.c:25-147): zero at first touch,
 * per face  val = 0.5*(var[p0]+var[p1]);  grad[p0] += n*val;  grad[p1] -= n*val,  scale by
 * 1/pvolume at last touch.  Here every own point belongs to exactly one tile; one thread owns
 * the point, walks its incident faces in the reference's single-thread order and keeps the 7x3
 * sums in registers: no atomics, no read of grad, one 168-byte row store per point.
 *
 * gg_tile_pipe_kernel (the production kernel): two CTAs per SM, each walking a chunk of consecutive tiles through
 * one shared-memory stage.  Per tile ONE elected thread issues the TMA bulk copies (cp.async.bulk global->shared,
 * mbarrier complete_tx): the tile blob (face normals read once, halo row list, ELL adjacency), the contiguous hvar
 * rows and volumes of the tile's own points; all threads gather the hvar rows of the tile's halo points with 8-byte
 * cp.async (LDGSTS) tracked by the same mbarrier.  The next tile is fetched as soon as the face walk of the
 * current one is over, the result rows leave through one TMA bulk store: HBM is only touched by asynchronous copies.
 *
 * Arithmetic modes
 *   EXACT = true : separate IEEE multiply and add in the reference's order.  The device holds
 *                  hvar = 0.5*var (halved once when var is uploaded): 0.5*(a+b) == 0.5*a + 0.5*b bit
 *                  for bit (power-of-two scaling commutes with rounding; subnormal inputs
 *                  excepted) -> bit-identical to the reference built without FMA, one thread.
 *   EXACT = false: fused multiply-add (more accurate, not bit-identical).
 */
#ifndef CFDP_GG_KERNELS_V4_CUH
#define CFDP_GG_KERNELS_V4_CUH

#include <cuda_runtime.h>
#include <stdint.h>
#include "common.h"

#define CFDP_MAX_HALO_POS_V4 1024  /* halo positions of a tile (shared-memory index list of the prefetcher) */
#define CFDP_MAX_CHUNK_V4 64      /* tiles per CTA */
#define CFDP_MAX_GATHER_PER_THREAD_V4 24 /* 8-byte cp.async a thread issues per tile for the halo gather (measured safe) */
#define CFDP_MAX_EXPORT_V4 256    /* export rows of a tile kept in shared memory (longer lists are read from global memory) */

namespace ggk4 {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
/* L2 eviction policies: the tile blobs and the gradient rows are touched once per iteration (evict first);
 * the hvar rows are read again by the neighbouring tiles' halo gathers (evict last) */
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void bulk_g2s_hint(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t policy)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void cp_async8(void *dst, const void *src)
{
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(void *dst, const void *src)
{
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(void *dst, const void *src)
{
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t *bar)
{
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ double flip_sign(double x, uint32_t signbit)
{
  return __hiloint2double(__double2hiint(x) ^ (int)signbit, __double2loint(x));
}

/* the face walk of one point for the equations [LO, LO+CNT): ell = this thread's ELL column */
template <bool EXACT, int LO, int CNT>
__device__ __forceinline__ void walk_faces(const uint32_t *__restrict__ ell, int npad, int maxdeg, const double *__restrict__ s_nrm,
                                           const double *__restrict__ s_hvar, const double (&hv)[CNT], double (&acc)[CNT * 3])
{
  uint32_t e_next = maxdeg > 0 ? ell[0] : CFDP_ADJ_PAD;
#pragma unroll 2
  for (int j = 0; j < maxdeg; j++) {
    const uint32_t e = e_next;                                    /* the adjacency entry is fetched one step ahead */
    e_next = j + 1 < maxdeg ? ell[(j + 1) * npad] : CFDP_ADJ_PAD;
    if (e == CFDP_ADJ_PAD) continue;
    const double *n = s_nrm + 3 * ((e >> 16) & 0x7FFFu);
    const double *w = s_hvar + NGRAD * (e & 0x7FFFu) + LO; /* bit 15: the neighbour is a ghost (flux_kernels.cuh) */
    const uint32_t sb = e & 0x80000000u;      /* this point is p1 of the face: grad[p1] -= n*val (gradients.c:101-105) */
    const double nx = flip_sign(n[0], sb), ny = flip_sign(n[1], sb), nz = flip_sign(n[2], sb);
#pragma unroll
    for (int q = 0; q < CNT; q++) {
      if (EXACT) {
        const double val = __dadd_rn(hv[q], w[q]);                     /* == 0.5*(var[p0]+var[p1]), gradients.c:77 */
        acc[3 * q + 0] = __dadd_rn(acc[3 * q + 0], __dmul_rn(nx, val));
        acc[3 * q + 1] = __dadd_rn(acc[3 * q + 1], __dmul_rn(ny, val));
        acc[3 * q + 2] = __dadd_rn(acc[3 * q + 2], __dmul_rn(nz, val));
      } else {
        const double val = hv[q] + w[q];
        acc[3 * q + 0] = fma(nx, val, acc[3 * q + 0]);
        acc[3 * q + 1] = fma(ny, val, acc[3 * q + 1]);
        acc[3 * q + 2] = fma(nz, val, acc[3 * q + 2]);
      }
    }
  }
}

/* one point, equations [LO, LO+CNT): walk, scale by 1/volume, park the partial row in the output staging */
template <bool EXACT, int LO, int CNT>
__device__ __forceinline__ void point_rows(int p, const uint32_t *__restrict__ ell0, int npad, int maxdeg, const double *__restrict__ s_nrm,
                                           const double *__restrict__ s_hvar, double inv_vol, double (&acc)[CNT * 3])
{
  double hv[CNT];
#pragma unroll
  for (int q = 0; q < CNT; q++) hv[q] = s_hvar[p * NGRAD + LO + q];
#pragma unroll
  for (int k = 0; k < CNT * 3; k++) acc[k] = 0.0;
  walk_faces<EXACT, LO, CNT>(ell0 + p, npad, maxdeg, s_nrm, s_hvar, hv, acc);
#pragma unroll
  for (int k = 0; k < CNT * 3; k++) acc[k] = __dmul_rn(acc[k], inv_vol);
}

/* shared-memory stage of a tile: [blob | output rows (aliased)][half-var rows][volumes], each part 128-byte aligned;
 * the offsets depend on the tile, stage_bytes is the largest footprint of any tile */
__host__ __device__ __forceinline__ uint32_t tile_var_off(uint32_t blob_bytes, uint32_t npts)
{
  const uint32_t out_bytes = npts * (NGRAD * 3 * 8);
  return ((blob_bytes > out_bytes ? blob_bytes : out_bytes) + 127u) & ~127u;
}
__host__ __device__ __forceinline__ uint32_t tile_pvol_off(uint32_t blob_bytes, uint32_t npts, uint32_t nhalo)
{
  return tile_var_off(blob_bytes, npts) + (((CFDP_HALO_BASE(npts) + nhalo) * (NGRAD * 8) + 127u) & ~127u);
}
__host__ __device__ __forceinline__ uint32_t tile_footprint(uint32_t blob_bytes, uint32_t npts, uint32_t nhalo)
{
  return tile_pvol_off(blob_bytes, npts, nhalo) + ((CFDP_HALO_BASE(npts) * 8 + 127u) & ~127u);
}
struct PipeLayout {
  uint32_t stage_bytes;
  int block_points;  /* threads per CTA (multiple of 32, >= largest tile) */
  unsigned long long *prof; /* optional: SM cycles of thread 0 summed over tiles: [0] wait for data, [1] face walk, [2] rest of the tile, [4] tiles */
  unsigned long long *progress; /* counts finished boundary tiles (tile index < nsignal): the early-send trigger the comm stream waits on */
  int nsignal;
  int split_roles; /* warp 0 drives the bulk copies instead of gathering */
  int tile_base;  /* global index of this launch's first tile */
  int nexport;    /* global tiles [0, nexport) write their export rows (fused pack); 0 = off */
  const uint32_t *exp_off, *exp_src, *exp_dst; /* per boundary tile: tile-local point -> grad row (bit 31 clear) or send-buffer slot (bit 31 set) */
  double *sendbuf;
};

__device__ __forceinline__ void bulk_s2g(void *gdst, const void *ssrc, uint32_t bytes, uint64_t policy)
{
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
               ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes), "l"(policy) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
/* the bulk stores counted by n are complete (wait_group): publish them with one release reduction */
__device__ __forceinline__ void signal_progress(unsigned long long *ctr, int n)
{
  asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(ctr), "l"((unsigned long long)n) : "memory");
}
__device__ __forceinline__ void bulk_wait_prev() { asm volatile("cp.async.bulk.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

/*
 * Two CTAs per SM, each owning one shared-memory stage and a chunk of consecutive tiles.  Timeline of a tile t:
 *   wait(mbarrier)            all of tile t has landed (bulk copies + halo gather); then the halo row list of tile
 *                             t+1 is prefetched (16-byte cp.async) and lands during the walk
 *   face walk                 one thread per point, 21 sums in registers
 *   S1 barrier                normals / adjacency / var of tile t are dead
 *   stage the rows of tile t  into the head of the stage (transposed through shared memory)
 *   S2 barrier
 *   TMA bulk store of the rows (one instruction); boundary tiles also write their export rows (fused pack)
 *   early fetch of tile t+1   while the TMA engine drains the staged rows: var rows, volumes, blob tail (bulk copies)
 *                             and the halo gather (8-byte cp.async): everything that does not overlap the staged rows
 *   head of tile t+1's blob   once the store has read shared memory.  While this CTA waits, the other CTA computes.
 * No thread ever blocks on a global load: HBM is touched only by asynchronous copies.
 */
template <bool EXACT>
__global__ void __launch_bounds__(CFDP_MAX_TILE_POINTS, 2)
gg_tile_pipe_kernel(const TileDesc *__restrict__ tiles, int ntiles, int chunk, const unsigned char *__restrict__ blob,
                    const double *__restrict__ hvar, const double *__restrict__ pvol, double *__restrict__ grad, PipeLayout L)
{
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full;
  __shared__ TileDesc s_tds[CFDP_MAX_CHUNK_V4];   /* descriptors of this CTA's tiles */
  __shared__ __align__(16) uint32_t s_hidx[CFDP_MAX_HALO_POS_V4]; /* halo row list of the tile being prefetched */
  __shared__ uint32_t s_exp[2 * CFDP_MAX_EXPORT_V4];               /* this tile's export list: sources, then destinations */
  __shared__ uint32_t s_exp_off[CFDP_MAX_CHUNK_V4 + 1];            /* export list bounds of this CTA's tiles */
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int t_begin = blockIdx.x * chunk;
  const int t_end = min(t_begin + chunk, ntiles);
  if (t_begin >= t_end) return;
  {
    const int nw = (t_end - t_begin) * (int)(sizeof(TileDesc) / 4);
    const uint32_t *g = reinterpret_cast<const uint32_t *>(tiles + t_begin);
    uint32_t *d = reinterpret_cast<uint32_t *>(s_tds);
    for (int i = tid; i < nw; i += nthr) d[i] = __ldg(g + i);
    for (int i = tid; i <= t_end - t_begin; i += nthr) {
      const int gt = L.tile_base + t_begin + i;
      s_exp_off[i] = gt <= L.nexport ? __ldg(L.exp_off + gt) : 0u; /* exp_off has nexport + 1 entries */
    }
  }
  if (tid == 0) {
    mbar_init(&full, (uint32_t)nthr + 1);
    fence_mbar_init();
  }
  unsigned char *const st = smem;
  const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();

  /* halo row list of tile t (part of its blob) -> s_hidx, asynchronously */
  auto stage_pf_index = [&](int t) {
    if (t < t_end) {
      const TileDesc pd = s_tds[t - t_begin];
      const uint4 *g = reinterpret_cast<const uint4 *>(blob + pd.blob_off() + pd.halo_off);
      const int n16 = (int)pd.nhalo >> 2; /* nhalo (positions) is a multiple of 16 */
      for (int i = tid; i < n16; i += nthr) cp_async16(&s_hidx[4 * i], g + i);
    }
    cp_async_commit();
  };
  /* thread 0: announce the bytes of tile t and start its bulk copies for blob bytes [lo, hi) (+ var rows and volumes) */
  auto bulk_part = [&](const TileDesc &pd, uint32_t lo, uint32_t hi, bool with_var, bool announce) {
    const uint32_t n_even = CFDP_HALO_BASE((uint32_t)pd.npts);
    const uint32_t nv = n_even * (NGRAD * 8), np = n_even * 8;
    if (announce) mbar_arrive_expect_tx(&full, pd.blob_bytes + nv + np);
    if (hi > lo) bulk_g2s_hint(st + lo, blob + pd.blob_off() + lo, hi - lo, &full, pol_stream);
    if (with_var) {
      bulk_g2s_hint(st + tile_var_off(pd.blob_bytes, pd.npts), hvar + (size_t)pd.row0 * NGRAD, nv, &full, pol_keep);
      bulk_g2s_hint(st + tile_pvol_off(pd.blob_bytes, pd.npts, pd.nhalo), pvol + pd.row0, np, &full, pol_stream);
    }
  };
  /* all threads: hvar rows of the halo points, consecutive lanes = consecutive words of a row.  (Moving a row as
   * 16-byte pieces needs the halo position to share the parity of the device row; measured slower: the parity
   * constraint costs more bank conflicts in the face walk than the shorter gather saves.) */
  auto gather_halo = [&](const TileDesc &pd, int first, int stride) { /* first < 0: this thread only arrives */
    double *vs = reinterpret_cast<double *>(st + tile_var_off(pd.blob_bytes, pd.npts)) + (size_t)CFDP_HALO_BASE((uint32_t)pd.npts) * NGRAD;
    const int nw = (int)pd.nhalo * NGRAD;
    for (int i = first >= 0 ? first : nw; i < nw; i += stride) {
      const int r = i / NGRAD;
      const uint32_t row = s_hidx[r];
      if (row != 0xFFFFFFFFu) cp_async8(vs + i, hvar + (size_t)row * NGRAD + (i - r * NGRAD));
    }
    /* the arrival is issued by whole, converged warps only: issued under divergence (lane 0 of warp 0 busy with the
     * bulk copies, or lanes leaving the loop above at different trips) the phase was observed to complete early */
    __syncwarp();
    cp_async_mbar_arrive_noinc(&full);
  };

  __syncthreads(); /* descriptors and mbarrier visible */
  stage_pf_index(t_begin);
  cp_async_wait_all();
  __syncthreads();
  {
    const TileDesc pd = s_tds[0];
    if (tid == 0) bulk_part(pd, 0, pd.blob_bytes, true, true);
    gather_halo(pd, tid, nthr);
  }
  __syncthreads(); /* s_hidx may be refilled */
  /* split roles (CFDP_SPLIT_ROLES=0 turns it off): warp 0 does not gather, its lane 0 drives the bulk store and the
   * bulk copies of the next tile, so that the head of the next blob is requested the moment the store has drained
   * the staged rows (4 % faster: 1.82 -> 1.75 ms on 16.8 M points) */
  const bool split_roles = nthr >= 128 && L.split_roles == 1;
  const int g_first = split_roles ? (tid >= 32 ? tid - 32 : -1) : tid, g_stride = split_roles ? nthr - 32 : nthr;

  int pending_sig = 0; /* thread 0: boundary tiles stored but not yet signalled */
  for (int t = t_begin, it = 0; t < t_end; ++t, ++it) {
    const bool has_next = t + 1 < t_end;
    const TileDesc td = s_tds[t - t_begin];
    const int npts = td.npts, nhalo = td.nhalo;
    double *s_nrm = reinterpret_cast<double *>(st);
    const double *s_hvar = reinterpret_cast<const double *>(st + tile_var_off(td.blob_bytes, td.npts));
    const double *s_pvol = reinterpret_cast<const double *>(st + tile_pvol_off(td.blob_bytes, td.npts, td.nhalo));
    const uint32_t *ell0 = reinterpret_cast<const uint32_t *>(st + td.halo_off + ((nhalo * 4 + 15) & ~15));
    long long c0 = 0, c1 = 0, c2 = 0;
    if (L.prof && tid == 0) c0 = clock64();
    mbar_wait(&full, (uint32_t)it & 1u);
    if (L.prof && tid == 0) c1 = clock64();
    /* the phase completes only after every thread has arrived, i.e. is done reading s_hidx for this tile's halo
     * gather: the list of the next tile may now be fetched; it lands during the face walk */
    stage_pf_index(t + 1);

    double acc[NGRAD * 3];
    const bool active = tid < npts;
    if (active) {
      const double inv_vol = __ddiv_rn(1.0, s_pvol[tid]);                 /* gradients.c:138 */
      point_rows<EXACT, 0, NGRAD>(tid, ell0, td.npad, td.maxdeg, s_nrm, s_hvar, inv_vol, acc);
    }
    cp_async_wait_all(); /* this thread's share of the next halo row list is in s_hidx */
    __syncthreads();     /* S1: normals, adjacency and var of this tile are dead */
    if (L.prof && tid == 0) c2 = clock64();

    /* the output rows are staged in [0, out_end) and leave through one bulk store; while the TMA engine reads them,
     * whatever of the next tile lives beyond out_end is fetched; the head of its blob follows when the read is done */
    const uint32_t out_rows = CFDP_HALO_BASE((uint32_t)npts);
    const uint32_t out_end = (out_rows * (NGRAD * 3 * 8) + 127u) & ~127u;
    TileDesc nd = td;
    bool early = false;
    if (has_next) {
      nd = s_tds[t + 1 - t_begin];
      early = tile_var_off(nd.blob_bytes, nd.npts) >= out_end;
    }
    /* export list of this tile (fused pack), fetched asynchronously while the rows are being staged */
    const int gt = L.tile_base + t;
    const uint32_t e0 = gt < L.nexport ? s_exp_off[t - t_begin] : 0u;
    const int nexp = gt < L.nexport ? (int)(s_exp_off[t - t_begin + 1] - e0) : 0;
    const bool exp_in_smem = nexp > 0 && nexp <= CFDP_MAX_EXPORT_V4;
    if (exp_in_smem) {
      for (int i = tid; i < nexp; i += nthr) {
        cp_async4(&s_exp[i], L.exp_src + e0 + i);
        cp_async4(&s_exp[CFDP_MAX_EXPORT_V4 + i], L.exp_dst + e0 + i);
      }
      cp_async_commit();
    }
    if (active) {
      double *o = s_nrm + tid * (NGRAD * 3);
#pragma unroll
      for (int k = 0; k < NGRAD * 3; k++) o[k] = acc[k];
    }
    fence_proxy_async(); /* the staged rows (generic proxy) become visible to the bulk store (async proxy) */
    cp_async_wait_all(); /* export list */
    __syncthreads();     /* S2 */
    long long q1 = 0, q2 = 0, q3 = 0;
    if (L.prof && tid == 0) q1 = clock64();
    if (tid == 0) {
      bulk_s2g(grad + (size_t)td.row0 * (NGRAD * 3), s_nrm, out_rows * (NGRAD * 3 * 8), pol_stream); /* rows beyond npts are alignment padding */
      bulk_commit();
    }
    if (nexp > 0) {
      /* fused pack (threads.c:187-249, :791-813): the rows of this tile that other domains need go straight from the
       * staged rows to the packed send buffer, or to the ghost rows of a domain hosted on this GPU */
      const int nw = nexp * (NGRAD * 3);
      for (int i = tid; i < nw; i += nthr) {
        const int r = i / (NGRAD * 3), c = i - r * (NGRAD * 3);
        const uint32_t src = exp_in_smem ? s_exp[r] : __ldg(L.exp_src + e0 + r);
        const uint32_t dst = exp_in_smem ? s_exp[CFDP_MAX_EXPORT_V4 + r] : __ldg(L.exp_dst + e0 + r);
        double *base = (dst & 0x80000000u) ? L.sendbuf : grad;
        base[(size_t)(dst & 0x7FFFFFFFu) * (NGRAD * 3) + c] = s_nrm[src * (NGRAD * 3) + c];
      }
      __syncthreads(); /* the staged rows have been read by every thread (and ordered before thread 0's release below) */
    }
    if (has_next && early) { /* runs while the bulk store drains the staged rows */
      if (tid == 0) bulk_part(nd, out_end < nd.blob_bytes ? out_end : nd.blob_bytes, nd.blob_bytes, true, true);
      gather_halo(nd, g_first, g_stride);
    }
    if (L.prof && tid == 0) q2 = clock64();
    if (tid == 0) {
      bulk_wait_read();      /* shared memory may be overwritten */
      if (L.prof) q3 = clock64();
      /* boundary tiles: their rows may be packed / copied by the exchange as soon as every boundary tile has retired
       * (the reference's finalised-send-point counters, threads.c:268-306).  A CTA reports its boundary tiles in
       * one go, when it retires or reaches its first interior tile: nobody waits for a write to reach global memory. */
      if (pending_sig && t >= L.nsignal) {   /* first interior tile of this CTA: flush the signals of its boundary tiles */
        bulk_wait_prev();    /* every store group but the one just committed is complete */
        signal_progress(L.progress, pending_sig);
        pending_sig = 0;
      }
      if (t < L.nsignal) pending_sig++;
      if (has_next) {
        if (early) bulk_part(nd, 0, out_end < nd.blob_bytes ? out_end : nd.blob_bytes, false, false);
        else bulk_part(nd, 0, nd.blob_bytes, true, true);
      }
    }
    if (has_next && !early) { /* rare (a much smaller tile follows): its var rows overlap the staged rows */
      __syncthreads();
      gather_halo(nd, tid, nthr);
      __syncthreads();
    }
    if (L.prof && tid == 0) {
      const long long c3 = clock64();
      atomicAdd(L.prof + 0, (unsigned long long)(c1 - c0)); atomicAdd(L.prof + 1, (unsigned long long)(c2 - c1));
      atomicAdd(L.prof + 2, (unsigned long long)(c3 - c2)); atomicAdd(L.prof + 4, 1ull);
      atomicAdd(L.prof + 5, (unsigned long long)(q1 - c2)); atomicAdd(L.prof + 6, (unsigned long long)(q2 - q1)); atomicAdd(L.prof + 7, (unsigned long long)(q3 - q2));
      /* [5] staging up to S2, [6] store issue + exports + early fetch, [7] wait for the store to have read shared memory */
    }
  }
  if (tid == 0 && pending_sig) {
    bulk_wait_all();
    signal_progress(L.progress, pending_sig);
  }
}

} // namespace ggk4
#endif
