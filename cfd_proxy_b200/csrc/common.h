/*
 * common.h -- internal declarations shared by the host code and the CUDA engine.
 * Not part of the C ABI (that is include/cfdp_b200.h).
 */
#ifndef CFDP_COMMON_H
#define CFDP_COMMON_H

#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <vector>
#include "cfdp_b200.h"

/* print-and-exit convention of the reference (error_handling.h:29-36) */
#define ASSERT(x...)                                                        \
  do {                                                                      \
    if (!(x)) {                                                             \
      fprintf(stderr, "Error: '%s' [%s:%i]\n", #x, __FILE__, __LINE__);     \
      exit(EXIT_FAILURE);                                                   \
    }                                                                       \
  } while (0)

#define CFDP_TILE_ALIGN 16      /* device rows per alignment unit: 16 rows * 56 B = 7 * 128 B */
#define CFDP_MAX_TILE_POINTS 256
#define CFDP_ADJ_PAD 0xFFFFFFFFu

/* one tile of the GPU face schedule: a group of own points of one domain whose incident
 * faces are processed by one thread block.  32 bytes. */
struct TileDesc {
  uint32_t row0;       /* first device row of the tile's points (global over the GPU) */
  uint16_t npts;       /* own points in the tile */
  uint16_t nhalo;      /* points outside the tile its faces reference */
  uint16_t nfaces;     /* face slots in the tile blob (multiple of 16, <= 32767) */
  uint16_t zslot;      /* a face slot no face uses: its normal is (0,0,0).  The production kernel turns padding entries of the
                        * adjacency into (neighbour = the point itself, face = zslot): a contribution of exactly +-0 */
  uint16_t maxdeg;     /* ELL depth: max incident faces of a tile point */
  uint16_t npad;       /* ELL row pitch (npts rounded up to 32) */
  uint32_t blob128;    /* offset of the tile blob in units of 128 bytes */
  uint32_t hrow0;      /* first row of the tile's block in the packed halo array (nhalo rows, see engine.cu) */
  uint32_t blob_bytes; /* size of the tile blob (multiple of 128) */
  uint32_t halo_off;   /* byte offset of the halo row list inside the blob (normals come first) */
#ifdef __CUDACC__
  __host__ __device__
#endif
  size_t blob_off() const { return (size_t)blob128 * 128; }
};
/* tile-local point index: [0,npts) own points of the tile, halo points from CFDP_HALO_BASE(npts) on
 * (even, so that the own var rows can be fetched as one 16-byte granular bulk copy) */
#define CFDP_HALO_BASE(npts) (((npts) + 1) & ~1)
/* gather mode (gg_tile_gather_kernel): var rows are 64-byte rows fetched by TMA tensor copies in boxes of 16 own rows and
 * groups of 4 halo rows, so the halo positions start at a multiple of 16 */
#define CFDP_HALO_BASE_G(npts) (((npts) + 15) & ~15)

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

/* tile blob: [normals nfaces*3 f64][pad16][halo rows nhalo u32][pad16][ELL maxdeg*npad u32][pad16] */
static inline size_t blob_halo_off(uint32_t nfaces) { return align_up((size_t)nfaces * 24, 16); }
static inline size_t blob_adj_off(uint32_t nfaces, uint32_t nhalo) { return blob_halo_off(nfaces) + align_up((size_t)nhalo * 4, 16); }
static inline size_t blob_size(uint32_t nfaces, uint32_t nhalo, uint32_t maxdeg, uint32_t npad)
{
  return align_up(blob_adj_off(nfaces, nhalo) + (size_t)maxdeg * npad * 4, 128);
}

struct DomainSchedule {
  int nown = 0, nall = 0;
  int nrows = 0;                        /* own rows incl. tile alignment padding + ghost rows */
  int ghost_row0 = 0;                   /* first ghost row (domain relative) */
  int ntiles = 0, nboundary = 0;
  long long nfaces_computed = 0;        /* faces with >= 1 own endpoint (rangelist.c:513-523) */
  long long tile_faces = 0, halo_refs = 0;
  std::vector<int> row_of_point;        /* [nall] domain-relative device row */
  std::vector<int> tile_row0;           /* [ntiles+1] */
  std::vector<int> tile_npts, tile_nfaces, tile_nhalo, tile_maxdeg, tile_is_boundary;
  std::vector<int> tile_nslots, tile_nhpos; /* shared-memory positions of face normals / halo rows (multiples of 16, >= the counts) */
  std::vector<int> tile_zslot;          /* per tile: an unused face slot (zero normal), see TileDesc::zslot */
  std::vector<uint64_t> tile_blob;      /* [ntiles+1] byte offsets into blob */
  std::vector<unsigned char> blob;      /* halo rows domain-relative until commit rebases them */
  std::vector<int> tile_face_ids;       /* concatenated original face ids in slot order */
  std::vector<long long> tile_face_off; /* [ntiles+1] */
  std::vector<int> tile_halo_pts;       /* concatenated original local point ids */
  std::vector<long long> tile_halo_off; /* [ntiles+1] */
  /* pseudo flux (flux.c): per tile a second, smaller blob in the same format -- only the adjacency entries that
   * contribute (this point is p1 of the face, or p0 with a ghost p1), the normals of those faces (every face once
   * per domain instead of once per incident tile) and the halo rows those entries reference */
  std::vector<unsigned char> fblob;
  std::vector<uint64_t> ftile_blob;     /* [ntiles+1] */
  std::vector<int> ftile_nfaces, ftile_nhalo, ftile_maxdeg;
  int max_nfaces = 0, max_nloc = 0;     /* per-tile maxima: faces, local points (own, even-padded, + halo) */
  size_t max_blob = 0;                  /* largest tile blob in bytes */
  long long lds_wavefronts_min = 0, lds_wavefronts_est = 0; /* face-walk shared-memory wavefronts: conflict free / estimated */
};

struct ScheduleOptions {
  int tile_points;     /* max own points per tile (<= CFDP_MAX_TILE_POINTS, multiple of 16) */
  int max_faces;       /* cap of face records per tile */
  int max_local;       /* cap of npts + nhalo */
  int order;           /* 0 = greedy graph growing, 1 = consecutive chunks of the file numbering */
  int bank_placement;  /* 1 = place face slots and halo rows by shared-memory bank (fewer conflicts), 0 = discovery order */
  int slack_slots, slack_halo; /* spare face slots / halo positions per tile for the bank placement */
  int flux_blob;       /* 1 = also build the pseudo-flux blobs */
  int refine_rounds;   /* local-search sweeps of the bank placement after the greedy pass (0 = greedy only) */
  int sort_in_tile;    /* 1 = points of a tile in ascending file numbering, 0 = in growth (BFS) order */
  int stage_budget;    /* bytes one tile may occupy in shared memory (blob + var rows + volumes); 0 = no limit */
  int gather;          /* 1 = schedule for gg_tile_gather_kernel: 64-byte var rows (volume in word 7), halo positions from align16(npts),
                        * halo rows placed by 16-byte bank group (class = position mod 8, groups = quarter warps) */
};

/* schedule.cpp */
void build_schedule(const solver_data *sd, const comm_data *cd, const ScheduleOptions &opt, DomainSchedule &out);

struct Domain {
  int id = -1;                 /* domain rank (cd->iProc) */
  comm_data *cd = nullptr;
  solver_data *sd = nullptr;
  bool comm_read = false, tables_done = false, threads_inited = false;
  DomainSchedule sch;
  long long rowbase = 0;       /* first device row of this domain */
  long long tile0_b = 0, tile0_i = 0; /* first boundary / interior tile in the global tile list */
};

/* engine.cu */
struct Engine;
Engine *engine_get(void);
int engine_proc_of_domain(int domain);
Domain *engine_find_domain(const void *cd_or_sd);
Domain *engine_domain_by_id(int id);
Domain *engine_register_domain(comm_data *cd, int id);
void engine_exchange_ints(const std::vector<int> &peer, const std::vector<const int *> &sbuf, const std::vector<int> &scount,
                          const std::vector<int *> &rbuf, const std::vector<int> &rcount);
void *engine_alloc_pinned(size_t bytes);
void engine_free_pinned(void *p);
int engine_num_hosted(void);
Domain *engine_hosted(int i);

#endif
