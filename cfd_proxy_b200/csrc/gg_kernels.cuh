/*
 * gg_kernels.cuh -- the Green-Gauss tile kernels (sm_100a).
 *
 * Replaces private_compute_gradients_gg (reference src/gradients.c:25-147): zero at first touch,
 * per face  val = 0.5*(var[p0]+var[p1]);  grad[p0] += n*val;  grad[p1] -= n*val,  scale by
 * 1/pvolume at last touch.  Here every own point belongs to exactly one tile; one thread owns
 * the point, walks its incident faces in the reference's single-thread order and keeps the 7x3
 * sums in registers: no atomics, no read of grad, one 168-byte row store per point.
 *
 * gg_tile_pipe_kernel (the production kernel): two CTAs per SM, each walking a chunk of consecutive tiles through
 * one shared-memory stage.  Per tile ONE elected thread issues the TMA bulk copies (cp.async.bulk global->shared,
 * mbarrier complete_tx): the tile blob (face normals read once, ELL adjacency), the contiguous hvar rows and volumes
 * of the tile's own points and the tile's PACKED halo rows (halo_pack_kernel writes them when var is uploaded; there
 * is no gather inside the kernel).  The next tile is fetched as soon as the face walk of the current one is over --
 * the part of it that must wait for the result rows to leave is asked into L2 beforehand (cp.async.bulk.prefetch.L2)
 * -- and the result rows leave through one TMA bulk store per warp: HBM is only touched by asynchronous copies.
 *
 * Arithmetic modes
 *   EXACT = true : separate IEEE multiply and add in the reference's order.  The device holds
 *                  hvar = 0.5*var (halved once when var is uploaded): 0.5*(a+b) == 0.5*a + 0.5*b bit
 *                  for bit (power-of-two scaling commutes with rounding; subnormal inputs
 *                  excepted) -> bit-identical to the reference built without FMA, one thread.
 *   EXACT = false: fused multiply-add (more accurate, not bit-identical).
 */
#ifndef CFDP_GG_KERNELS_CUH
#define CFDP_GG_KERNELS_CUH

#include <cuda_runtime.h>
#include <stdint.h>
#include "common.h"

#define CFDP_MAX_HALO_POS 1024  /* bound on the halo positions of a tile (sanity check at commit) */
#define CFDP_MAX_CHUNK 64      /* tiles per CTA */
#define CFDP_MAX_EXPORT 256    /* export rows of a tile kept in shared memory (longer lists are read from global memory) */

namespace ggk {

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
 * the hvar rows are read again by halo_pack_kernel only (evict last is kept for uploads that repack at once) */
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

/* ------------------------------------------------------------------------------------------------------------------
 * Shared-memory stage of a tile (gg_tile_pipe_kernel), stage_bytes = the same for every tile of a launch:
 *
 *   [0, blob_bytes)                          the tile blob: normals | halo row list | ELL adjacency
 *   [0, n_even * 168)                        the result rows, staged here once the face walk is over (aliases the blob)
 *   [hvar_off, hvar_off + (n_even+nhalo)*56) half-var rows: own points, then the halo positions   } end aligned: the next
 *   [pvol_off, stage_bytes)                  volumes of the own points                             } tile's rows never
 *                                                                                                   } overlap staged rows
 * The staged rows of warp w are bytes [5376 w, 5376 (w+1)) ("zone" w: 32 rows of 168 bytes): every warp stages, stores
 * (one TMA bulk store per warp) and refills its own zone without a CTA-wide barrier.
 * ---------------------------------------------------------------------------------------------------------------- */
#define CFDP_ROW_BYTES (NGRAD * 3 * 8)       /* 168 */
#define CFDP_ZONE_BYTES (32 * CFDP_ROW_BYTES) /* 5376 */

__host__ __device__ __forceinline__ uint32_t stage_hvar_bytes(uint32_t npts, uint32_t nhalo)
{
  return ((CFDP_HALO_BASE(npts) + nhalo) * (NGRAD * 8) + 127u) & ~127u;
}
__host__ __device__ __forceinline__ uint32_t stage_pvol_bytes(uint32_t npts) { return (CFDP_HALO_BASE(npts) * 8 + 127u) & ~127u; }
__host__ __device__ __forceinline__ uint32_t stage_pvol_off(uint32_t stage_bytes, uint32_t npts) { return stage_bytes - stage_pvol_bytes(npts); }
__host__ __device__ __forceinline__ uint32_t stage_hvar_off(uint32_t stage_bytes, uint32_t npts, uint32_t nhalo)
{
  return stage_pvol_off(stage_bytes, npts) - stage_hvar_bytes(npts, nhalo);
}
/* bytes one tile needs: blob (or the staged rows, whichever is larger) + var rows + volumes */
__host__ __device__ __forceinline__ uint32_t tile_footprint(uint32_t blob_bytes, uint32_t npts, uint32_t nhalo)
{
  const uint32_t out_bytes = CFDP_HALO_BASE(npts) * CFDP_ROW_BYTES;
  return (((blob_bytes > out_bytes ? blob_bytes : out_bytes) + 127u) & ~127u) + stage_hvar_bytes(npts, nhalo) + stage_pvol_bytes(npts);
}

#define CFDP_EXP_BASES 10 /* export destinations: 0 = grad of this GPU, 1 = packed send buffer, 2 + k = grad of peer GPU k (CUDA IPC mapping) */
struct PipeLayout {
  uint32_t stage_bytes;
  int cstride, istride, maxcount;  /* CTA b walks tiles b*cstride + i*istride, i < maxcount (contiguous chunks: chunk,1,chunk; interleaved: 1,grid,inf) */
  unsigned long long *prof;        /* optional: SM cycles of thread 0 summed over tiles: [0] wait for data, [1] face walk, [2] rest, [4] tiles */
  unsigned long long *progress;    /* counts finished boundary tiles (tile index < nsignal): the early-send trigger the comm stream waits on */
  int nsignal;
  int variant;                     /* CFDP_VARIANT: 1 (default) = L2 prefetch of the next tile's late-fetched blob head at the start of the
                                    * face walk, 2 = of all of the next tile, 0 = none */
  int spread_m, spread_nb;         /* > 1: boundary tile q < spread_nb is walked at position q * spread_m (see tile_of) */
  int tile_base;                   /* global index of this launch's first tile */
  int nexport;                     /* global tiles [0, nexport) write their export rows (fused pack / direct halo stores); 0 = off */
  const uint32_t *exp_off, *exp_src, *exp_dst; /* per boundary tile: (tile-local point | destination array << 16) -> row of that array */
  double *exp_base[CFDP_EXP_BASES];
  /* direct halo stores into peer memory: per boundary tile the (peer, rows) pairs it completes; the peer's arrival counter
   * is bumped by the number of rows (threads.c:268-306: per-partner counters of finalised send points) */
  const uint32_t *sig_off, *sig_ent;          /* ent = peer | rows << 4 */
  unsigned long long *sig_flag[CFDP_EXP_BASES - 2];
};

__device__ __forceinline__ void bulk_s2g(void *gdst, uint32_t ssrc, uint32_t bytes, uint64_t policy)
{
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
               ::"l"(gdst), "r"(ssrc), "r"(bytes), "l"(policy) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_prev() { asm volatile("cp.async.bulk.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
/* the stores counted by n are complete: publish them with one release reduction */
__device__ __forceinline__ void signal_progress(unsigned long long *ctr, int n)
{
  asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(ctr), "l"((unsigned long long)n) : "memory");
}
__device__ __forceinline__ void signal_peer(unsigned long long *ctr, unsigned long long n)
{
  asm volatile("red.release.sys.global.add.u64 [%0], %1;" ::"l"(ctr), "l"(n) : "memory");
}
/* shared-memory accesses of the face walk by 32-bit shared address: the compiler neither rematerialises the
 * (cluster-relative) base of the dynamic shared memory in the loop nor reorders them across the barriers */
__device__ __forceinline__ double lds_f64(uint32_t a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts_f64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void bulk_g2s_a(uint32_t sdst, const void *src, uint32_t bytes, uint32_t bar, uint64_t policy)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(sdst), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void *src, uint32_t bytes)
{
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async8_a(uint32_t sdst, const void *src)
{
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sdst), "l"(src) : "memory");
}

struct FaceOps { double n[3]; double w[NGRAD]; };
/* operands of one adjacency entry: the face normal and the neighbour's half-var row */
__device__ __forceinline__ void load_face_ops(uint32_t e, uint32_t a_nrm, uint32_t a_hv, FaceOps &o)
{
  const uint32_t an = a_nrm + 24u * ((e >> 16) & 0x7FFFu), aw = a_hv + (NGRAD * 8u) * (e & 0x7FFFu);
#pragma unroll
  for (int c = 0; c < 3; c++) o.n[c] = lds_f64(an + 8u * c);
#pragma unroll
  for (int q = 0; q < NGRAD; q++) o.w[q] = lds_f64(aw + 8u * q);
}
template <bool EXACT>
__device__ __forceinline__ void face_update(const FaceOps &o, uint32_t e, const double (&hv)[NGRAD], double (&acc)[NGRAD * 3])
{
  const uint32_t sb = e & 0x80000000u;        /* this point is p1 of the face: grad[p1] -= n*val (gradients.c:101-105) */
  const double nx = flip_sign(o.n[0], sb), ny = flip_sign(o.n[1], sb), nz = flip_sign(o.n[2], sb);
#pragma unroll
  for (int q = 0; q < NGRAD; q++) {
    if (EXACT) {
      const double val = __dadd_rn(hv[q], o.w[q]);                      /* == 0.5*(var[p0]+var[p1]), gradients.c:77 */
      acc[3 * q + 0] = __dadd_rn(acc[3 * q + 0], __dmul_rn(nx, val));
      acc[3 * q + 1] = __dadd_rn(acc[3 * q + 1], __dmul_rn(ny, val));
      acc[3 * q + 2] = __dadd_rn(acc[3 * q + 2], __dmul_rn(nz, val));
    } else {
      const double val = hv[q] + o.w[q];
      acc[3 * q + 0] = fma(nx, val, acc[3 * q + 0]);
      acc[3 * q + 1] = fma(ny, val, acc[3 * q + 1]);
      acc[3 * q + 2] = fma(nz, val, acc[3 * q + 2]);
    }
  }
}

/*
 * Packed halo rows.  The half-var rows of a tile's halo points (end points of its faces outside the tile) are not
 * gathered by the gradient kernel: a row-wise gather through the load/store unit (8-byte cp.async) cost 20 % of the kernel
 * (measured: 1.81 -> 1.45 ms on 16.8 M points without it), because every warp-level copy of scattered 56-byte rows takes
 * about one L2 round trip to issue while the other CTA of the SM keeps the unit busy with its face walk.  Instead,
 * whenever var is uploaded this kernel writes, per tile, a contiguous copy of its halo rows in the order of the tile's
 * shared-memory halo positions (halo_pack_kernel); the gradient kernel fetches the block with ONE TMA bulk copy.
 * Price: 65 bytes per point of extra DRAM reads per iteration (the copies are not shared between tiles).
 */
__global__ void __launch_bounds__(256)
halo_pack_kernel(const TileDesc *__restrict__ tiles, int ntiles, const unsigned char *__restrict__ blob,
                 const double *__restrict__ hvar, double *__restrict__ hhalo)
{
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const TileDesc td = tiles[t];
    const uint32_t *hrows = reinterpret_cast<const uint32_t *>(blob + td.blob_off() + td.halo_off);
    double *dst = hhalo + (size_t)td.hrow0 * NGRAD;
    const int nw = (int)td.nhalo * NGRAD;
    for (int k = threadIdx.x; k < nw; k += blockDim.x) {
      const int r = k / NGRAD;
      const uint32_t row = __ldg(hrows + r);
      dst[k] = row != 0xFFFFFFFFu ? __ldg(hvar + (size_t)row * NGRAD + (k - r * NGRAD)) : 0.0; /* unused positions: zeros */
    }
  }
}

/*
 * The production kernel.  Two CTAs per SM, each owning one shared-memory stage and walking a sequence of tiles.
 * Timeline of tile t in one CTA:
 *   B0   barrier (the descriptor prefetched during the previous walk is visible) + mbarrier wait: the bulk copies of
 *        tile t (blob, own var rows, packed halo rows, volumes) have landed.
 *   walk one thread per own point: its adjacency column in the reference's face order, 21 sums in registers; the
 *        operands of step j+1 are loaded while step j is computed; padding entries are turned into exact zeros
 *        (zero normal, the point itself as neighbour), so the loop has no data-dependent branch.
 *   S1   barrier: normals / adjacency / var rows of tile t are dead.
 *        early fetch of tile t+1 (thread 0, four bulk copies): everything that does not overlap the staged rows --
 *        blob tail, own var rows, halo rows, volumes -- is requested NOW, before the rows are staged.
 *   per warp, no CTA barrier: stage the 32 rows of the warp (zone w), fence.proxy.async, lane 0 issues one TMA bulk
 *        store for the zone, waits until the store has READ the zone and refills it at once with the bytes of tile
 *        t+1's blob that live there (late fetch).  Boundary tiles also write their export rows (fused pack, or direct
 *        stores into the ghost rows of a peer GPU) between two barriers, and bump the arrival counters.
 * No thread ever touches global memory with a load: HBM is read by TMA bulk copies only.
 */
template <bool EXACT, int CTAS>
__global__ void __launch_bounds__(CTAS == 4 ? 128 : CTAS == 3 ? 160 : CFDP_MAX_TILE_POINTS, CTAS)
gg_tile_pipe_kernel(const TileDesc *__restrict__ tiles, int ntiles, const unsigned char *__restrict__ blob,
                    const double *__restrict__ hvar, const double *__restrict__ hhalo, const double *__restrict__ pvol,
                    double *__restrict__ grad, PipeLayout L)
{
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full;
  __shared__ __align__(16) TileDesc s_desc[4];                  /* ring: descriptors of tiles i, i+1, i+2 of this CTA */
  __shared__ uint32_t s_eoff[4][2];                              /* ring: export list bounds of those tiles */
  __shared__ uint32_t s_exp[2 * CFDP_MAX_EXPORT];                /* this tile's export list: sources, then destinations */
  __shared__ uint32_t s_soff[4][2];                              /* ring: bounds of those tiles' peer-signal entries (direct halo stores) */
  __shared__ uint32_t s_sig[CFDP_EXP_BASES - 2];                 /* this tile's peer-signal entries */
  __shared__ unsigned long long s_sigacc[CFDP_EXP_BASES - 2];    /* rows stored into each peer's memory and not yet signalled (thread 0 only) */
  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5;
  /* this CTA's tiles: t(i) = first + i*istride, i < count */
  const long long first = (long long)blockIdx.x * L.cstride;
  if (first >= ntiles) return;
  int count = (int)((ntiles - first + L.istride - 1) / L.istride);
  if (count > L.maxcount) count = L.maxcount;
  const uint32_t sbase = smem_u32(smem), bar = smem_u32(&full);
  const uint64_t pol_stream = l2_policy_evict_first();
  /* position k of the walk -> tile.  Normally the identity (boundary tiles come first in the tile list: early send).  With direct
   * halo stores nothing waits for the boundary as a whole, and a burst of boundary tiles at the start of the kernel only congests
   * NVLink with their small stores: the spread_nb boundary tiles are then dealt out evenly, one at the head of every run of
   * spread_m positions (CFDP_DIRECT_SPREAD) */
  auto tile_of = [&](int i) {
    const long long k = first + (long long)i * L.istride;
    if (L.spread_m <= 1) return (int)k;
    const long long q = k / L.spread_m, r = k - q * L.spread_m;
    if (q >= L.spread_nb) return (int)k;                              /* past the mixed part: the remaining interior tiles in order */
    return r == 0 ? (int)q : (int)(L.spread_nb + q * (L.spread_m - 1) + (r - 1));
  };

  /* descriptor (and export bounds) of this CTA's tile i -> ring slot i & 3, asynchronously */
  auto prefetch_desc = [&](int i) {
    if (i < count) {
      const int t = tile_of(i);
      if (tid < 8) cp_async4(reinterpret_cast<uint32_t *>(&s_desc[i & 3]) + tid, reinterpret_cast<const uint32_t *>(tiles + t) + tid);
      else if (tid < 10) {
        const int gt = L.tile_base + t + (tid - 8);
        if (gt <= L.nexport && L.nexport > 0) cp_async4(&s_eoff[i & 3][tid - 8], L.exp_off + gt); /* exp_off has nexport + 1 entries */
      } else if (tid < 12) {
        const int gt = L.tile_base + t + (tid - 10);
        if (gt <= L.nexport && L.nexport > 0 && L.sig_off) cp_async4(&s_soff[i & 3][tid - 10], L.sig_off + gt);
      }
    }
  };
  /* bytes [lo, hi) of tile pd's blob -> the same offsets of the stage, minus the halo row list [halo_off, ell_off): only
   * halo_pack_kernel reads that list, the face walk never does */
  auto blob_fetch = [&](const TileDesc &pd, uint32_t lo, uint32_t hi) {
    const uint32_t h0 = pd.halo_off, h1 = pd.halo_off + (((uint32_t)pd.nhalo * 4u + 15u) & ~15u);
    const unsigned char *src = blob + pd.blob_off();
    if (lo < h0 && lo < hi) bulk_g2s_a(sbase + lo, src + lo, (hi < h0 ? hi : h0) - lo, bar, pol_stream);
    const uint32_t l2 = lo > h1 ? lo : h1;
    if (l2 < hi) bulk_g2s_a(sbase + l2, src + l2, hi - l2, bar, pol_stream);
  };
  /* thread 0: bulk copies of tile pd: blob bytes [lo, blob_bytes), own var rows, packed halo rows, volumes;
   * announces ALL bytes of the tile (the rest of the blob follows from the warps' late fetches).  Issuing these after
   * the stores, or spreading them over four warps, measured the same within 1.5 % (profiles/README.md) */
  auto bulk_early = [&](const TileDesc &pd, uint32_t lo) {
    const uint32_t n_even = CFDP_HALO_BASE((uint32_t)pd.npts);
    const uint32_t nv = n_even * (NGRAD * 8), nh = (uint32_t)pd.nhalo * (NGRAD * 8), np = n_even * 8;
    const uint32_t a_hv = sbase + stage_hvar_off(L.stage_bytes, pd.npts, pd.nhalo);
    const uint32_t hole = ((uint32_t)pd.nhalo * 4u + 15u) & ~15u;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(pd.blob_bytes - hole + nv + nh + np) : "memory");
    blob_fetch(pd, lo, pd.blob_bytes);
    bulk_g2s_a(a_hv, hvar + (size_t)pd.row0 * NGRAD, nv, bar, pol_stream);
    if (nh) bulk_g2s_a(a_hv + nv, hhalo + (size_t)pd.hrow0 * NGRAD, nh, bar, pol_stream);
    bulk_g2s_a(sbase + stage_pvol_off(L.stage_bytes, pd.npts), pvol + pd.row0, np, bar, pol_stream);
  };

  /* prologue: descriptors of tiles 0, 1 (2 follows asynchronously), all of tile 0 */
  if (tid == 0) {
    mbar_init(&full, 1);
    fence_mbar_init();
#pragma unroll
    for (int k = 0; k < CFDP_EXP_BASES - 2; k++) s_sigacc[k] = 0ull;
  }
  prefetch_desc(0); prefetch_desc(1);
  cp_async_commit(); cp_async_wait_all();
  __syncthreads();
  if (tid == 0) bulk_early(s_desc[0], 0);

  int pending_sig = 0; /* boundary tiles of this CTA stored but not yet reported (uniform over the CTA) */
  for (int i = 0; i < count; ++i) {
    const int t = tile_of(i);
    const bool has_next = i + 1 < count;
    long long c0 = 0, c1 = 0, c2 = 0;
    if (L.prof && tid == 0) c0 = clock64();
    cp_async_wait_all();
    __syncthreads();                    /* B0: the prefetched descriptor of tile i+1 is visible to every thread */
    mbar_wait(&full, (uint32_t)i & 1u); /* the bulk copies have landed */
    if (L.prof && tid == 0) c1 = clock64();
    const TileDesc td = s_desc[i & 3];
    const int npts = td.npts;
    prefetch_desc(i + 2);
    /* export list of this tile (boundary tiles): fetched now, it lands during the face walk */
    const int gt = L.tile_base + t;
    const uint32_t e0 = gt < L.nexport ? s_eoff[i & 3][0] : 0u;
    const int nexp = gt < L.nexport ? (int)(s_eoff[i & 3][1] - e0) : 0;
    const bool exp_in_smem = nexp > 0 && nexp <= CFDP_MAX_EXPORT;
    const uint32_t g0 = (nexp > 0 && L.sig_off) ? s_soff[i & 3][0] : 0u;
    const int nsig = (nexp > 0 && L.sig_off) ? (int)(s_soff[i & 3][1] - g0) : 0;   /* <= one entry per peer */
    /* NOTE: threads 32.. fetch the entries, so direct halo stores need blocks of at least 64 threads (tiles of more than 32
     * points); smaller tiles are only used by single-GPU tests, where sig_off is null */
    if (tid >= 32 && tid < 32 + nsig) cp_async4(&s_sig[tid - 32], L.sig_ent + g0 + (tid - 32));
    if (exp_in_smem) {
      for (int k = tid; k < nexp; k += nthr) {
        cp_async4(&s_exp[k], L.exp_src + e0 + k);
        cp_async4(&s_exp[CFDP_MAX_EXPORT + k], L.exp_dst + e0 + k);
      }
    }
    cp_async_commit();
    if (L.variant && has_next && tid == 0) {
      /* the bytes of the next tile that can only be fetched late (the head of its blob, once this tile's result rows have
       * left the stage) are asked into L2 now: their bulk copies then cost an L2 round trip instead of a DRAM one
       * (+5.5 .. 7 %).  variant 2 asks for all of the next tile: measured slower than the head alone */
      const TileDesc &pd = s_desc[(i + 1) & 3];
      const unsigned char *src = blob + pd.blob_off();
      const uint32_t oc = CFDP_HALO_BASE((uint32_t)npts) * CFDP_ROW_BYTES;
      if (L.variant == 2) {
        const uint32_t ne = CFDP_HALO_BASE((uint32_t)pd.npts);
        bulk_prefetch_l2(src, (pd.blob_bytes + 15u) & ~15u);
        bulk_prefetch_l2(hvar + (size_t)pd.row0 * NGRAD, ne * (NGRAD * 8));
        if (pd.nhalo) bulk_prefetch_l2(hhalo + (size_t)pd.hrow0 * NGRAD, (uint32_t)pd.nhalo * (NGRAD * 8));
        bulk_prefetch_l2(pvol + pd.row0, ne * 8);
      } else {
        bulk_prefetch_l2(src, ((oc < pd.blob_bytes ? oc : pd.blob_bytes) + 15u) & ~15u);
      }
    }

    double acc[NGRAD * 3];
#pragma unroll
    for (int k = 0; k < NGRAD * 3; k++) acc[k] = 0.0;
    const bool active = tid < npts;
    if (active) {
      const uint32_t a_hv = sbase + stage_hvar_off(L.stage_bytes, td.npts, td.nhalo);
      const uint32_t a_ell = sbase + td.halo_off + (((uint32_t)td.nhalo * 4u + 15u) & ~15u) + 4u * tid;
      const uint32_t pitch = 4u * td.npad;
      const int maxdeg = td.maxdeg;
      const uint32_t pad_e = (uint32_t)tid | ((uint32_t)td.zslot << 16); /* the point itself across a zero normal */
      double hv[NGRAD];
#pragma unroll
      for (int q = 0; q < NGRAD; q++) hv[q] = lds_f64(a_hv + (NGRAD * 8u) * tid + 8u * q);
      const double vol = lds_f64(sbase + stage_pvol_off(L.stage_bytes, td.npts) + 8u * tid);
      if (maxdeg > 0) {
        /* adjacency entry of step j: raw word (steps beyond the column re-read the last row), then padding entries
         * and steps beyond the column become the point itself across the zero normal: a contribution of exactly +-0 */
        auto raw = [&](int j) { return lds_u32(a_ell + pitch * (uint32_t)min(j, maxdeg - 1)); };
        auto fix = [&](uint32_t r, int j) { return (j < maxdeg && r != CFDP_ADJ_PAD) ? r : pad_e; };
        FaceOps A, B;
        uint32_t ea = fix(raw(0), 0), eb = fix(raw(1), 1);
        load_face_ops(ea, sbase, a_hv, A);
        for (int j = 0; j < maxdeg; j += 2) {
          const uint32_t ra = raw(j + 2), rb = raw(j + 3);   /* two steps ahead: address arithmetic off the critical path */
          load_face_ops(eb, sbase, a_hv, B);
          face_update<EXACT>(A, ea, hv, acc);
          ea = fix(ra, j + 2);
          load_face_ops(ea, sbase, a_hv, A);
          face_update<EXACT>(B, eb, hv, acc);
          eb = fix(rb, j + 3);
        }
      }
      const double inv_vol = __ddiv_rn(1.0, vol);                        /* gradients.c:138 */
#pragma unroll
      for (int k = 0; k < NGRAD * 3; k++) acc[k] = __dmul_rn(acc[k], inv_vol);
    }
    cp_async_wait_all(); /* this thread's share of the export list (boundary tiles) and of the descriptor prefetch */
    __syncthreads();     /* S1: normals, adjacency and var of this tile are dead; the export list is complete */
    if (L.prof && tid == 0) c2 = clock64();

    const uint32_t n_even = CFDP_HALO_BASE((uint32_t)npts);
    const uint32_t out_cover = n_even * CFDP_ROW_BYTES;    /* bytes [0, out_cover) hold the staged rows */
    TileDesc nd = td;
    if (has_next) { /* early fetch of the next tile: runs under the staging / store of this one */
      nd = s_desc[(i + 1) & 3];
      if (tid == 0) bulk_early(nd, out_cover < nd.blob_bytes ? out_cover : nd.blob_bytes);
    }
    long long q1 = 0, q2 = 0, q3 = 0;
    if (L.prof && tid == 0) q1 = clock64();

    /* stage this warp's rows; rows npts .. n_even-1 are alignment padding of the tile (zeros) */
    const uint32_t rows_w = (uint32_t)(32 * warp) < n_even ? min(32u, n_even - 32u * warp) : 0u;
    if ((uint32_t)tid < n_even) {
      const uint32_t o = sbase + CFDP_ROW_BYTES * (uint32_t)tid;
#pragma unroll
      for (int k = 0; k < NGRAD * 3; k++) sts_f64(o + 8u * k, acc[k]);
    }
    if (rows_w) {
      fence_proxy_async(); /* the staged rows (generic proxy) become visible to the bulk store (async proxy) */
      __syncwarp();
      if (lane == 0) {
        bulk_s2g(grad + ((size_t)td.row0 + 32u * warp) * (NGRAD * 3), sbase + CFDP_ZONE_BYTES * (uint32_t)warp, rows_w * CFDP_ROW_BYTES, pol_stream);
        bulk_commit();
      }
    }
    if (L.prof && tid == 0) q2 = clock64();
    if (nexp > 0) {
      /* fused pack (threads.c:187-249, :791-813) / direct halo stores: the rows of this tile that other domains need go
       * straight from the staged rows to their consumers: the packed send buffer, the ghost rows of a domain hosted on
       * this GPU, or the ghost rows of a domain on a peer GPU (CUDA IPC mapping, stores over NVLink).  (Letting every
       * warp export the rows of its own zone without the two barriers measured slower: 5.4 % against 1.7 % of a
       * 16.8 M-point iteration.) */
      __syncthreads();
      /* a 168-byte row leaves as ten 16-byte stores and one 8-byte store (at its head when the destination row is odd:
       * rows are 8-byte aligned, every second one 16-byte aligned); one store per thread and step */
      const int nw = nexp * 11;
      for (int k = tid; k < nw; k += nthr) {
        const int r = k / 11, c = k - r * 11;
        const uint32_t sk = exp_in_smem ? s_exp[r] : __ldg(L.exp_src + e0 + r);     /* tile-local point | destination array << 16 */
        const uint32_t dst = exp_in_smem ? s_exp[CFDP_MAX_EXPORT + r] : __ldg(L.exp_dst + e0 + r);
        double *g = L.exp_base[sk >> 16] + (size_t)dst * (NGRAD * 3);
        const uint32_t a = sbase + CFDP_ROW_BYTES * (sk & 0xFFFFu);
        const uint32_t odd = dst & 1u;
        if (c == 10) {        /* the 8-byte piece: word 0 of an odd row, word 20 of an even one */
          const uint32_t w = odd ? 0u : 20u;
          g[w] = lds_f64(a + 8u * w);
        } else {
          const uint32_t w = 2u * (uint32_t)c + odd;
          const double2 v = make_double2(lds_f64(a + 8u * w), lds_f64(a + 8u * w + 8u));
          *reinterpret_cast<double2 *>(g + w) = v;
        }
      }
      __syncthreads(); /* the staged rows have been read by every thread; the stores are ordered before thread 0's releases */
      if (tid == 0) /* rows this tile stores into peer memory: signalled in one go per CTA, below */
        for (int q = 0; q < nsig; q++) s_sigacc[s_sig[q] & 15u] += (unsigned long long)(s_sig[q] >> 4);
    }
    if (rows_w && lane == 0) {
      bulk_wait_read();      /* the zone may be overwritten: late fetch of the next tile's bytes that live in it */
      if (L.prof && tid == 0) q3 = clock64();
      if (has_next) {
        const uint32_t lo = CFDP_ZONE_BYTES * (uint32_t)warp, hi = min(min(lo + rows_w * CFDP_ROW_BYTES, out_cover), (uint32_t)nd.blob_bytes);
        blob_fetch(nd, lo, hi);
      }
    }
    /* direct halo stores: once this CTA has stored the rows of its last exporting tile it bumps the arrival counters of
     * the peers (red.release.sys after a system fence: the stores of all threads were ordered before thread 0 by the
     * barrier that ended the export).  One fence per CTA, not per tile: it costs an NVLink round trip, and it comes
     * after the late fetch so that the next tile is not held up by it (threads.c:268-306: per partner counters). */
    if (L.sig_off && tid == 0 && gt < L.nexport && (!has_next || L.tile_base + tile_of(i + 1) >= L.nexport)) {
      bool any = false;
#pragma unroll
      for (int k = 0; k < CFDP_EXP_BASES - 2; k++) any = any || s_sigacc[k] != 0ull;
      if (any) {
        __threadfence_system();
#pragma unroll
        for (int k = 0; k < CFDP_EXP_BASES - 2; k++)
          if (s_sigacc[k]) { signal_peer(L.sig_flag[k], s_sigacc[k]); s_sigacc[k] = 0ull; }
      }
    }
    /* boundary tiles: their rows may be consumed by the exchange as soon as every boundary tile has retired (the
     * reference's finalised-send-point counters, threads.c:268-306).  A CTA reports its boundary tiles in one go,
     * when it reaches its first interior tile or retires: nobody waits for a write to reach global memory. */
    if (pending_sig && t >= L.nsignal) {
      if (lane == 0) { /* every store group of this warp but the one just committed (if any) is complete */
        if (rows_w) bulk_wait_prev(); else bulk_wait_all();
        fence_proxy_async_all();
      }
      __syncthreads();
      if (tid == 0) signal_progress(L.progress, pending_sig);
      pending_sig = 0;
    }
    if (t < L.nsignal) pending_sig++;
    if (L.prof && tid == 0) {
      const long long c3 = clock64();
      atomicAdd(L.prof + 0, (unsigned long long)(c1 - c0)); atomicAdd(L.prof + 1, (unsigned long long)(c2 - c1));
      atomicAdd(L.prof + 2, (unsigned long long)(c3 - c2)); atomicAdd(L.prof + 4, 1ull);
      /* [5] early fetch issue, [6] staging + store issue, [7] exports + wait until the store has read the zone */
      atomicAdd(L.prof + 5, (unsigned long long)(q1 - c2)); atomicAdd(L.prof + 6, (unsigned long long)(q2 - q1)); atomicAdd(L.prof + 7, (unsigned long long)(q3 - q2));
    }
  }
  if (pending_sig) {
    if (lane == 0) { bulk_wait_all(); fence_proxy_async_all(); }
    __syncthreads();
    if (tid == 0) signal_progress(L.progress, pending_sig);
  }
}

/* v1: one tile per CTA, synchronous staging (kept as a second, independent implementation) */
template <bool EXACT>
__global__ void __launch_bounds__(CFDP_MAX_TILE_POINTS, 2)
gg_tile_kernel(const TileDesc *__restrict__ tiles, const unsigned char *__restrict__ blob,
               const double *__restrict__ var /* 0.5 * var */, const double *__restrict__ pvol, double *__restrict__ grad,
               int region0_doubles)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *s_r0 = reinterpret_cast<double *>(smem_raw);  /* normals, later the output rows */
  double *s_var = s_r0 + region0_doubles;               /* [npts(even) + nhalo][7] */
  const TileDesc td = tiles[blockIdx.x];
  const unsigned char *tb = blob + td.blob_off();
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int npts = td.npts, nhalo = td.nhalo, nfaces = td.nfaces;
  const int n_even = CFDP_HALO_BASE(npts);

  { /* normals */
    const double2 *g = reinterpret_cast<const double2 *>(tb);
    double2 *s = reinterpret_cast<double2 *>(s_r0);
    const int n2 = (nfaces * 3 + 1) >> 1;
    for (int i = tid; i < n2; i += nthr) s[i] = __ldg(g + i);
  }
  { /* var rows of the tile's own points: contiguous, 16-byte aligned (row0 % 16 == 0) */
    const double2 *g = reinterpret_cast<const double2 *>(var + (size_t)td.row0 * NGRAD);
    double2 *s = reinterpret_cast<double2 *>(s_var);
    const int n2 = (n_even * NGRAD) >> 1;
    for (int i = tid; i < n2; i += nthr) s[i] = __ldg(g + i);
  }
  { /* var rows of the tile's halo points */
    const uint32_t *hrows = reinterpret_cast<const uint32_t *>(tb + td.halo_off);
    double *s = s_var + n_even * NGRAD;
    const int n = nhalo * NGRAD;
    for (int i = tid; i < n; i += nthr) {
      const int r = i / NGRAD, c = i - r * NGRAD;
      const uint32_t row = __ldg(hrows + r);
      if (row != 0xFFFFFFFFu) s[i] = __ldg(var + (size_t)row * NGRAD + c);
    }
  }
  __syncthreads();

  double acc[NGRAD * 3];
#pragma unroll
  for (int k = 0; k < NGRAD * 3; k++) acc[k] = 0.0;
  if (tid < npts) {
    double hv[NGRAD];
#pragma unroll
    for (int q = 0; q < NGRAD; q++) hv[q] = s_var[tid * NGRAD + q];
    const uint32_t *ell = reinterpret_cast<const uint32_t *>(tb + td.halo_off + ((nhalo * 4 + 15) & ~15)) + tid;
    walk_faces<EXACT, 0, NGRAD>(ell, td.npad, td.maxdeg, s_r0, s_var, hv, acc);   /* adjacency straight from global memory */
    const double tmp = __ddiv_rn(1.0, __ldg(pvol + td.row0 + tid));
#pragma unroll
    for (int k = 0; k < NGRAD * 3; k++) acc[k] = __dmul_rn(acc[k], tmp);
  }
  __syncthreads();
  if (tid < npts) {
#pragma unroll
    for (int k = 0; k < NGRAD * 3; k++) s_r0[tid * (NGRAD * 3) + k] = acc[k];
  }
  __syncthreads();
  {
    double *gout = grad + (size_t)td.row0 * (NGRAD * 3);
    const int n = npts * NGRAD * 3, n2 = n >> 1;
    double2 *g2 = reinterpret_cast<double2 *>(gout);
    const double2 *s2 = reinterpret_cast<const double2 *>(s_r0);
    for (int i = tid; i < n2; i += nthr) g2[i] = s2[i];
    if ((n & 1) && tid == 0) gout[n - 1] = s_r0[n - 1];
  }
}

} // namespace ggk
#endif
