/*
 * schedule.cpp -- the GPU face schedule, built once at setup from the same partition.
 *
 * Replaces the reference's CPU schedule: thread domains (src/rangelist.c:320-398), the
 * per-thread halo-first face sort and <=96-face colours with ftype 1/2/3
 * (src/rangelist.c:500-764), and the first/last-touch point lists
 * (src/points_of_color.c:26-287).  What is preserved:
 *   - only own points are written, faces with both ends in the outer halo are dropped
 *     (rangelist.c:513-523),
 *   - a face cut by a tile boundary is duplicated and each copy writes only its own end:
 *     the GPU analogue of ftype 1 / ftype 2 faces (rangelist.c:575-607, :719-736),
 *   - tiles that hold inner-halo (send) points come first, so their gradients can be packed
 *     and shipped while interior tiles compute (early send, threads.c:253-346),
 *   - every own point is zero-initialised and scaled by 1/volume exactly once: it lives in
 *     exactly one tile whose thread block owns its whole sum (eval.c:126-231 invariants).
 *
 * A tile = up to `tile_points` own points.  Its blob holds the normals of every face
 * incident to a tile point (read once per tile), the device rows of the non-tile end points
 * (halo of the tile) and a point-centric ELL adjacency: entry = neighbour's tile-local
 * index | ghost << 15 | face slot << 16 | sign << 31 (ghost: the neighbour is an addpoint of the domain; only the
 * pseudo-flux kernel looks at it).  A point's entries are sorted by the reference's
 * single-thread face order (ttype, p1, p0) (rangelist.c:567-608, util.c:113-136), so the
 * device sum runs over the same addends in the same order as the reference with one thread.
 *
 * Tiles are grown greedily over the own-point graph (BFS blobs, boundary first): the mesh
 * files carry no coordinates, and their numbering need not have any locality.
 */
#include <string.h>
#include <algorithm>
#include <vector>
#include <omp.h>
#include "common.h"

namespace {

struct AdjEntry { int face; int nbr; }; /* nbr >= 0: this point is p0 of face; encoded sign in face's top use */

struct Csr {
  std::vector<long long> off; /* [nown+1] */
  std::vector<int> face;      /* face id */
  std::vector<int> nbr;       /* other endpoint (local point id) */
  std::vector<unsigned char> sign; /* 0: point is p0 (+=), 1: point is p1 (-=) */
};

void build_csr(const solver_data *sd, Csr &g, long long &nfaces_computed)
{
  const int nown = sd->nownpoints, nf = sd->nfaces;
  g.off.assign((size_t)nown + 1, 0);
  long long nfc = 0;
  for (int f = 0; f < nf; f++) {
    const int p0 = sd->fpoint[f][0], p1 = sd->fpoint[f][1];
    ASSERT(p0 >= 0 && p0 < sd->nallpoints && p1 >= 0 && p1 < sd->nallpoints);
    if (p0 >= nown && p1 >= nown) continue;
    nfc++;
    if (p0 < nown) g.off[(size_t)p0 + 1]++;
    if (p1 < nown) g.off[(size_t)p1 + 1]++;
  }
  nfaces_computed = nfc;
  for (int p = 0; p < nown; p++) g.off[(size_t)p + 1] += g.off[p];
  const long long ne = g.off[nown];
  g.face.resize((size_t)ne); g.nbr.resize((size_t)ne); g.sign.resize((size_t)ne);
  std::vector<long long> cur(g.off.begin(), g.off.end() - 1);
  for (int f = 0; f < nf; f++) {
    const int p0 = sd->fpoint[f][0], p1 = sd->fpoint[f][1];
    if (p0 >= nown && p1 >= nown) continue;
    if (p0 < nown) { long long e = cur[p0]++; g.face[e] = f; g.nbr[e] = p1; g.sign[e] = 0; }
    if (p1 < nown) { long long e = cur[p1]++; g.face[e] = f; g.nbr[e] = p0; g.sign[e] = 1; }
  }
}

struct TileBuilder {
  const Csr &g; const ScheduleOptions &opt; int nown, nall;
  std::vector<int> tile_of;      /* [nown] */
  std::vector<int> halo_stamp;   /* [nall] */
  std::vector<int> queued_stamp; /* [nown] */
  std::vector<int> pts;          /* points in tile order, concatenated */
  std::vector<long long> tile_pt_off;
  /* current tile state */
  int cur = -1, np = 0, nf = 0, nh = 0, md = 0;

  TileBuilder(const Csr &g_, const ScheduleOptions &o, int nown_, int nall_) : g(g_), opt(o), nown(nown_), nall(nall_)
  {
    tile_of.assign((size_t)nown, -1); halo_stamp.assign((size_t)nall, -1); queued_stamp.assign((size_t)nown, -1);
    tile_pt_off.push_back(0);
  }
  void open() { cur = (int)tile_pt_off.size() - 1; np = nf = nh = md = 0; }
  /* shared-memory footprint of a tile: blob (normals, halo rows, ELL) + var rows + volumes */
  size_t footprint(int np_, int nf_, int nh_, int md_) const
  {
    const uint32_t npad = (uint32_t)align_up((size_t)np_, 32);
    nf_ = (int)align_up((size_t)nf_ + 1 + (size_t)opt.slack_slots, 16); nh_ = (int)align_up((size_t)nh_ + (size_t)opt.slack_halo, 16);
    if (opt.gather)
      return align_up(std::max(blob_size((uint32_t)nf_, (uint32_t)nh_, (uint32_t)md_, npad), (size_t)CFDP_HALO_BASE(np_) * CFDP_DIM2 * 8), 1024) +
             (size_t)(CFDP_HALO_BASE_G(np_) + nh_) * 64;
    return align_up(std::max(blob_size((uint32_t)nf_, (uint32_t)nh_, (uint32_t)md_, npad), (size_t)np_ * CFDP_DIM2 * 8), 128) +
           align_up((size_t)(CFDP_HALO_BASE(np_) + nh_) * NGRAD * 8, 128) + align_up((size_t)CFDP_HALO_BASE(np_) * 8, 128);
  }
  void close() { tile_pt_off.push_back((long long)pts.size()); cur = -1; }
  /* try to add own point p to the open tile; false when a cap would be exceeded */
  bool add(int p)
  {
    int in_tile = 0, new_halo = 0;
    const long long b = g.off[p], e = g.off[(size_t)p + 1];
    for (long long i = b; i < e; i++) {
      const int q = g.nbr[i];
      if (q < nown && tile_of[q] == cur) in_tile++;
      else if (halo_stamp[q] != cur) new_halo++; /* duplicates inside one point's list are rare; over-count is safe */
    }
    const int was_halo = halo_stamp[p] == cur ? 1 : 0;
    const int nf2 = nf + (int)(e - b) - in_tile;
    const int nloc2 = np + 1 + nh + new_halo - was_halo;
    const int md2 = std::max(md, (int)(e - b));
    if (np > 0 && (np + 1 > opt.tile_points || nf2 > opt.max_faces || nloc2 > opt.max_local)) return false;
    if (np > 0 && opt.stage_budget > 0 && footprint(np + 1, nf2, nloc2 - (np + 1), md2) > (size_t)opt.stage_budget) return false;
    ASSERT((int)(e - b) <= opt.max_faces && (int)(e - b) + 1 <= opt.max_local); /* a single point must fit */
    tile_of[p] = cur; pts.push_back(p); np++; nf = nf2; md = md2;
    nh -= was_halo;
    for (long long i = b; i < e; i++) {
      const int q = g.nbr[i];
      if (q < nown && tile_of[q] == cur) continue;
      if (halo_stamp[q] != cur) { halo_stamp[q] = cur; nh++; }
    }
    return true;
  }
};

} // namespace

static double wall(void) { return omp_get_wtime(); }
void build_schedule(const solver_data *sd, const comm_data *cd, const ScheduleOptions &opt, DomainSchedule &out)
{
  const bool prof = getenv("CFDP_SCHED_PROF") != nullptr;
  double t_prev = wall();
  auto lap = [&](const char *what) { if (prof) { const double t = wall(); fprintf(stderr, "  schedule[%d pts] %-28s %.2f s\n", sd->nownpoints, what, t - t_prev); t_prev = t; } };
  const int nown = sd->nownpoints, nall = sd->nallpoints;
  ASSERT(nown > 0 && nall >= nown);
  ASSERT(opt.tile_points >= 16 && opt.tile_points <= CFDP_MAX_TILE_POINTS && opt.tile_points % 16 == 0);
  ASSERT(opt.max_faces <= 32767 && opt.max_local <= 65534);
  out = DomainSchedule();
  out.nown = nown; out.nall = nall;

  /* halo type: 1 own, 2 inner halo (listed in some sendindex), 3 outer halo (rangelist.c:118-148) */
  std::vector<unsigned char> is_send((size_t)nown, 0);
  if (cd && cd->ndomains > 1)
    for (int i = 0; i < cd->ncommdomains; i++) {
      const int k = cd->commpartner[i];
      for (int j = 0; j < cd->sendcount[k]; j++) {
        const int p = cd->sendindex[k][j];
        ASSERT(p >= 0 && p < nown);
        is_send[p] = 1;
      }
    }

  Csr g;
  build_csr(sd, g, out.nfaces_computed);

  lap("csr");
  /* ---- 1. group own points into tiles ---- */
  TileBuilder tb(g, opt, nown, nall);
  if (opt.order == 1) {
    tb.open();
    for (int p = 0; p < nown; p++)
      if (!tb.add(p)) { tb.close(); tb.open(); ASSERT(tb.add(p)); }
    tb.close();
  } else {
    std::vector<int> fifoA, fifoB, q; /* seeds: send points / others; q: BFS queue of the open tile */
    size_t headA = 0, headB = 0;
    std::vector<int> send_list;
    for (int p = 0; p < nown; p++) if (is_send[p]) send_list.push_back(p);
    size_t scanA = 0; int scanB = 0;
    for (int phase = 0; phase < 2; phase++) {
      for (;;) {
        /* seed */
        int seed = -1;
        if (phase == 0) {
          while (headA < fifoA.size() && tb.tile_of[fifoA[headA]] >= 0) headA++;
          if (headA < fifoA.size()) seed = fifoA[headA++];
          else { while (scanA < send_list.size() && tb.tile_of[send_list[scanA]] >= 0) scanA++; if (scanA < send_list.size()) seed = send_list[scanA++]; }
        } else {
          while (headB < fifoB.size() && tb.tile_of[fifoB[headB]] >= 0) headB++;
          if (headB < fifoB.size()) seed = fifoB[headB++];
          else { while (scanB < nown && tb.tile_of[scanB] >= 0) scanB++; if (scanB < nown) seed = scanB++; }
        }
        if (seed < 0) break;
        if (tb.cur < 0) tb.open();
        q.clear(); size_t qh = 0;
        q.push_back(seed); tb.queued_stamp[seed] = tb.cur;
        bool full = false;
        while (qh < q.size()) {
          const int p = q[qh];
          if (tb.tile_of[p] >= 0) { qh++; continue; }
          if (!tb.add(p)) { full = true; break; }
          qh++;
          for (long long i = g.off[p]; i < g.off[(size_t)p + 1]; i++) {
            const int r = g.nbr[i];
            if (r < nown && tb.tile_of[r] < 0 && tb.queued_stamp[r] != tb.cur) { tb.queued_stamp[r] = tb.cur; q.push_back(r); }
          }
        }
        if (full || tb.np >= opt.tile_points) {
          /* the unvisited frontier seeds the following tiles, so they grow next to this one */
          for (size_t i = qh; i < q.size(); i++) {
            const int r = q[i];
            if (tb.tile_of[r] >= 0) continue;
            if (is_send[r]) fifoA.push_back(r); else fifoB.push_back(r);
          }
          tb.close();
        }
        /* else: component exhausted, keep filling the same tile from the next seed */
      }
      if (tb.cur >= 0 && tb.np > 0) tb.close(); /* boundary and interior tiles are never mixed */
    }
  }
  const int ntiles = (int)tb.tile_pt_off.size() - 1;
  ASSERT((long long)tb.pts.size() == nown);
  /* inside a tile keep the file's numbering: where it has structure (neighbour id = own id + const), the
   * lanes of a warp then walk neighbouring rows of shared memory and bank conflicts drop */
  if (opt.sort_in_tile) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int t = 0; t < ntiles; t++) std::sort(tb.pts.begin() + tb.tile_pt_off[t], tb.pts.begin() + tb.tile_pt_off[(size_t)t + 1]);
  }

  lap("tiling");
  /* ---- 1b. own rows by shared-memory bank.  A tile point is a lane of the face walk AND a var row other lanes
   * gather.  Permuting the points inside a block of 16 consecutive positions keeps every half-warp's membership (so
   * the sets of rows read together do not change) but lets each point pick the bank-pair class (position mod 16)
   * that collides least with the other rows of the groups it is gathered in.  Halo rows and face slots are placed
   * afterwards against these fixed own rows (pass B). */
  if (opt.bank_placement && !opt.gather) { /* gather mode: quarter-warp groups, a 16-block permutation would change their membership */
    const int nthr0 = omp_get_max_threads();
    std::vector<std::vector<int>> lmap0((size_t)nthr0);
#pragma omp parallel
    {
      std::vector<int> &lmap = lmap0[omp_get_thread_num()];
      lmap.assign((size_t)nall, -1);
      struct Key { int tt, p1, p0, face, nb; };
      std::vector<Key> keys;
      std::vector<int> nbr_of, deg, goff, glist, fill, newP, rcls;
      std::vector<unsigned char> cnt, gmax;
#pragma omp for schedule(dynamic, 16)
      for (int t = 0; t < ntiles; t++) {
        int *P = &tb.pts[tb.tile_pt_off[t]];
        const int n = (int)(tb.tile_pt_off[(size_t)t + 1] - tb.tile_pt_off[t]);
        if (n <= 16) continue;
        int md = 0;
        for (int i = 0; i < n; i++) { lmap[P[i]] = i; md = std::max(md, (int)(g.off[(size_t)P[i] + 1] - g.off[P[i]])); }
        nbr_of.assign((size_t)n * md, -1); deg.assign((size_t)n, 0);
        for (int i = 0; i < n; i++) {
          const int p = P[i];
          keys.clear();
          for (long long a = g.off[p]; a < g.off[(size_t)p + 1]; a++) {
            const int q = g.nbr[a];
            const int p0 = g.sign[a] ? q : p, p1 = g.sign[a] ? p : q;
            const int h0 = p0 >= nown ? 3 : (is_send[p0] ? 2 : 1), h1 = p1 >= nown ? 3 : (is_send[p1] ? 2 : 1);
            keys.push_back(Key{((h0 == 2 || h1 == 2) ? 0 : 3) + (h0 == 3 ? 0 : (h1 == 3 ? 1 : 2)), p1, p0, g.face[a], lmap[q]});
          }
          std::sort(keys.begin(), keys.end(), [](const Key &x, const Key &y) {
            if (x.tt != y.tt) return x.tt < y.tt;
            if (x.p1 != y.p1) return x.p1 < y.p1;
            if (x.p0 != y.p0) return x.p0 < y.p0;
            return x.face < y.face;
          });
          deg[i] = (int)keys.size();
          for (size_t j = 0; j < keys.size(); j++) nbr_of[(size_t)i * md + j] = keys[j].nb;
        }
        /* row r -> distinct groups (half-warp of the reader, step) it is read in */
        goff.assign((size_t)n + 1, 0);
        for (int pass = 0; pass < 2; pass++) {
          if (pass == 1) { for (int r = 0; r < n; r++) goff[(size_t)r + 1] += goff[r]; glist.assign((size_t)goff[n], -1); fill.assign(goff.begin(), goff.end() - 1); }
          for (int i = 0; i < n; i++)
            for (int j = 0; j < deg[i]; j++) {
              const int r = nbr_of[(size_t)i * md + j];
              if (r < 0) continue;
              const int gi = (i / 16) * md + j;
              if (pass == 0) goff[(size_t)r + 1]++;
              else {
                bool dup = false;
                for (int x = goff[r]; x < fill[r]; x++) if (glist[x] == gi) dup = true;
                if (!dup) glist[fill[r]++] = gi;
              }
            }
        }
        const int ngrp = ((n + 15) / 16) * md;
        if (opt.refine_rounds > 0) {
          /* hill climbing from the file order: two points of a 16-block trade places when that lowers the sum over
           * groups of the fullest class; never worse than the file order, so the result is always kept */
          cnt.assign((size_t)ngrp * 16, 0); gmax.assign((size_t)ngrp, 0);
          rcls.resize((size_t)n);
          for (int r = 0; r < n; r++) {
            rcls[r] = r & 15;
            for (int x = goff[r]; x < goff[(size_t)r + 1]; x++) if (glist[x] >= 0) cnt[(size_t)glist[x] * 16 + (r & 15)]++;
          }
          auto remax = [&](int gi) { const unsigned char *cc = &cnt[(size_t)gi * 16]; int mx = 0; for (int x = 0; x < 16; x++) mx = std::max(mx, (int)cc[x]); return mx; };
          for (int gi = 0; gi < ngrp; gi++) gmax[gi] = (unsigned char)remax(gi);
          auto shift = [&](int r, int from, int to) {
            for (int x = goff[r]; x < goff[(size_t)r + 1]; x++) if (glist[x] >= 0) { cnt[(size_t)glist[x] * 16 + from]--; cnt[(size_t)glist[x] * 16 + to]++; }
          };
          auto touched_delta = [&](int ra, int rb) { /* objective change over the groups of ra and rb, from the stored gmax */
            int d = 0;
            for (int x = goff[ra]; x < goff[(size_t)ra + 1]; x++) if (glist[x] >= 0) d += remax(glist[x]) - gmax[glist[x]];
            for (int x = goff[rb]; x < goff[(size_t)rb + 1]; x++) {
              const int gi = glist[x];
              if (gi < 0) continue;
              bool both = false;
              for (int y = goff[ra]; y < goff[(size_t)ra + 1]; y++) if (glist[y] == gi) both = true;
              if (!both) d += remax(gi) - gmax[gi];
            }
            return d;
          };
          for (int round = 0; round < opt.refine_rounds; round++) {
            int improved = 0;
            for (int b0 = 0; b0 < n; b0 += 16) {
              const int m = std::min(16, n - b0);
              for (int ra = b0; ra < b0 + m; ra++) {
                bool critical = false;
                for (int x = goff[ra]; x < goff[(size_t)ra + 1] && !critical; x++) {
                  const int gi = glist[x];
                  critical = gi >= 0 && gmax[gi] > 1 && cnt[(size_t)gi * 16 + rcls[ra]] == gmax[gi];
                }
                if (!critical) continue;
                int best_d = 0, best_rb = -1;
                for (int rb = b0; rb < b0 + m; rb++) {
                  if (rb == ra) continue;
                  const int ca = rcls[ra], cb = rcls[rb];
                  shift(ra, ca, cb); shift(rb, cb, ca);
                  const int d = touched_delta(ra, rb);
                  shift(ra, cb, ca); shift(rb, ca, cb);
                  if (d < best_d) { best_d = d; best_rb = rb; }
                }
                if (best_rb < 0) continue;
                const int ca = rcls[ra], cb = rcls[best_rb];
                shift(ra, ca, cb); shift(best_rb, cb, ca);
                rcls[ra] = cb; rcls[best_rb] = ca;
                for (int x = goff[ra]; x < goff[(size_t)ra + 1]; x++) if (glist[x] >= 0) gmax[glist[x]] = (unsigned char)remax(glist[x]);
                for (int x = goff[best_rb]; x < goff[(size_t)best_rb + 1]; x++) if (glist[x] >= 0) gmax[glist[x]] = (unsigned char)remax(glist[x]);
                improved++;
              }
            }
            if (!improved) break;
          }
          newP.assign((size_t)n, -1);
          for (int r = 0; r < n; r++) { const int np_ = (r & ~15) + rcls[r]; ASSERT(np_ < n && newP[np_] < 0); newP[np_] = P[r]; }
          for (int i = 0; i < n; i++) lmap[P[i]] = -1;
          for (int i = 0; i < n; i++) P[i] = newP[i];
          continue;
        }
        cnt.assign((size_t)ngrp * 16, 0); gmax.assign((size_t)ngrp, 0);
        newP.assign((size_t)n, -1);
        for (int b0 = 0; b0 < n; b0 += 16) {
          const int m = std::min(16, n - b0);
          unsigned used = 0;
          for (int i = b0; i < b0 + m; i++) {
            int best = -1, best_cost = 1 << 30;
            for (int c0 = 0; c0 < m; c0++) {
              const int c = (c0 + i) % m;
              if (used & (1u << c)) continue;
              int cost = 0;
              for (int x = goff[i]; x < goff[(size_t)i + 1]; x++) {
                const int gi = glist[x];
                if (gi < 0) continue;
                const int cc = cnt[(size_t)gi * 16 + c];
                cost += (cc + 1 > gmax[gi] ? 4 : 0) + cc;
              }
              if (cost < best_cost) { best_cost = cost; best = c; }
            }
            used |= 1u << best;
            newP[(size_t)b0 + best] = P[i];
            for (int x = goff[i]; x < goff[(size_t)i + 1]; x++) {
              const int gi = glist[x];
              if (gi < 0) continue;
              unsigned char &cc = cnt[(size_t)gi * 16 + best];
              cc++;
              if (cc > gmax[gi]) gmax[gi] = cc;
            }
          }
        }
        for (int i = 0; i < n; i++) lmap[P[i]] = -1;
        /* keep the permutation only where it beats the file order (structured numberings are already good) */
        long long after = 0, before = 0;
        for (int gi = 0; gi < ngrp; gi++) after += gmax[gi];
        cnt.assign((size_t)ngrp * 16, 0); gmax.assign((size_t)ngrp, 0);
        for (int r = 0; r < n; r++)
          for (int x = goff[r]; x < goff[(size_t)r + 1]; x++) {
            const int gi = glist[x];
            if (gi < 0) continue;
            unsigned char &cc = cnt[(size_t)gi * 16 + (r & 15)];
            cc++;
            if (cc > gmax[gi]) gmax[gi] = cc;
          }
        for (int gi = 0; gi < ngrp; gi++) before += gmax[gi];
        if (after * 100 < before * 85)
          for (int i = 0; i < n; i++) { ASSERT(newP[i] >= 0); P[i] = newP[i]; }
      }
    }
  }

  lap("own-row placement (1b)");
  /* ---- 2. boundary tiles first, rows ---- */
  std::vector<int> tile_bnd((size_t)ntiles, 0), order((size_t)ntiles);
  for (int t = 0; t < ntiles; t++)
    for (long long i = tb.tile_pt_off[t]; i < tb.tile_pt_off[(size_t)t + 1]; i++)
      if (is_send[tb.pts[i]]) { tile_bnd[t] = 1; break; }
  int nb = 0;
  for (int t = 0; t < ntiles; t++) if (tile_bnd[t]) order[nb++] = t;
  { int k = nb; for (int t = 0; t < ntiles; t++) if (!tile_bnd[t]) order[k++] = t; }
  out.ntiles = ntiles; out.nboundary = nb;
  out.tile_row0.assign((size_t)ntiles + 1, 0);
  out.tile_npts.resize(ntiles); out.tile_nfaces.resize(ntiles); out.tile_nhalo.resize(ntiles);
  out.tile_maxdeg.resize(ntiles); out.tile_is_boundary.resize(ntiles);
  out.row_of_point.assign((size_t)nall, -1);
  std::vector<long long> pt_off((size_t)ntiles + 1, 0); /* into tb.pts, new tile order */
  int row = 0;
  for (int k = 0; k < ntiles; k++) {
    const int t = order[k];
    const int n = (int)(tb.tile_pt_off[(size_t)t + 1] - tb.tile_pt_off[t]);
    out.tile_row0[k] = row; out.tile_npts[k] = n; out.tile_is_boundary[k] = tile_bnd[t];
    pt_off[k] = tb.tile_pt_off[t];
    for (int i = 0; i < n; i++) out.row_of_point[tb.pts[tb.tile_pt_off[t] + i]] = row + i;
    row += (int)align_up((size_t)n, CFDP_TILE_ALIGN);
  }
  out.tile_row0[ntiles] = row;
  out.ghost_row0 = row;
  for (int p = nown; p < nall; p++) out.row_of_point[p] = row + (p - nown);
  out.nrows = (int)align_up((size_t)row + (size_t)(nall - nown), CFDP_TILE_ALIGN);
  std::vector<int> tile_of_new((size_t)nown); /* new tile index of each own point */
  for (int k = 0; k < ntiles; k++)
    for (int i = 0; i < out.tile_npts[k]; i++) tile_of_new[tb.pts[pt_off[k] + i]] = k;

  lap("rows");
  /* ---- 3. per tile: face ids, halo points ---- */
  /* eslot[e] for adjacency entry e: tile-local id of its face inside the tile of the entry's point
   * (ids in discovery order; the shared-memory slot is chosen later) */
  std::vector<int> eslot(g.face.size(), -1);
  std::vector<int> tnh((size_t)ntiles, 0), tmaxdeg((size_t)ntiles, 0);
  out.tile_nslots.assign((size_t)ntiles, 0); out.tile_nhpos.assign((size_t)ntiles, 0); out.tile_zslot.assign((size_t)ntiles, 0);
  const int nthreads = omp_get_max_threads();
  std::vector<std::vector<int>> lmap_t((size_t)nthreads);
  out.tile_face_off.assign((size_t)ntiles + 1, 0); out.tile_halo_off.assign((size_t)ntiles + 1, 0);
  /* pass A: count faces / halo points per tile */
#pragma omp parallel
  {
    std::vector<int> &lmap = lmap_t[omp_get_thread_num()];
    lmap.assign((size_t)nall, -1);
    std::vector<int> halo_local;
#pragma omp for schedule(dynamic, 16)
    for (int k = 0; k < ntiles; k++) {
      const int n = out.tile_npts[k];
      const int *P = &tb.pts[pt_off[k]];
      for (int i = 0; i < n; i++) lmap[P[i]] = i;
      int nf = 0, nh = 0, md = 0;
      halo_local.clear();
      for (int i = 0; i < n; i++) {
        const int p = P[i];
        const long long b = g.off[p], e = g.off[(size_t)p + 1];
        md = std::max(md, (int)(e - b));
        for (long long a = b; a < e; a++) {
          const int q = g.nbr[a];
          const bool q_in = q < nown && lmap[q] >= 0 && lmap[q] < n;
          if (!q_in && lmap[q] < 0) { lmap[q] = n + nh; nh++; halo_local.push_back(q); }
          /* the p0 side numbers the face; the p1 side numbers it only when p0 is outside the tile */
          if (g.sign[a] == 0 || !q_in) eslot[a] = nf++;
        }
      }
      for (int i = 0; i < n; i++) lmap[P[i]] = -1;
      for (int q : halo_local) lmap[q] = -1;
      out.tile_nfaces[k] = nf; tnh[k] = nh; tmaxdeg[k] = md;
      /* shared-memory positions: 16 residue classes of equal size, so that slots / rows can be placed by bank */
      /* a little slack lets the bank placement avoid forced collisions when a class fills up */
      /* + 1: at least one slot stays unused, its normal is zero (TileDesc::zslot) */
      out.tile_nslots[k] = (int)align_up((size_t)nf + 1 + (size_t)opt.slack_slots, 16); out.tile_nhpos[k] = (int)align_up((size_t)nh + (size_t)opt.slack_halo, 16);
      ASSERT(out.tile_nslots[k] <= 32767 && n + 1 + out.tile_nhpos[k] <= 65534);
    }
  }
  for (int k = 0; k < ntiles; k++) {
    out.tile_nhalo[k] = tnh[k]; out.tile_maxdeg[k] = tmaxdeg[k];
    out.tile_face_off[(size_t)k + 1] = out.tile_face_off[k] + out.tile_nfaces[k];
    out.tile_halo_off[(size_t)k + 1] = out.tile_halo_off[k] + tnh[k];
    out.max_nfaces = std::max(out.max_nfaces, out.tile_nslots[k]);
    out.max_nloc = std::max(out.max_nloc, (opt.gather ? CFDP_HALO_BASE_G(out.tile_npts[k]) : CFDP_HALO_BASE(out.tile_npts[k])) + out.tile_nhpos[k]);
  }
  out.tile_faces = out.tile_face_off[ntiles]; out.halo_refs = out.tile_halo_off[ntiles];
  out.tile_face_ids.resize((size_t)out.tile_faces); out.tile_halo_pts.resize((size_t)out.halo_refs);
  out.tile_blob.assign((size_t)ntiles + 1, 0);
  for (int k = 0; k < ntiles; k++) {
    const uint32_t npad = (uint32_t)align_up((size_t)out.tile_npts[k], 32);
    const size_t bs = blob_size((uint32_t)out.tile_nslots[k], (uint32_t)out.tile_nhpos[k], (uint32_t)tmaxdeg[k], npad);
    out.tile_blob[(size_t)k + 1] = out.tile_blob[k] + bs;
    out.max_blob = std::max(out.max_blob, bs);
  }
  out.blob.assign((size_t)out.tile_blob[ntiles], 0);

  lap("pass A (counts)");
  /* pass B: order every point's faces, place face slots and halo rows by shared-memory bank, emit blobs */
  const bool place_by_bank = opt.bank_placement != 0;
  std::vector<std::vector<unsigned char>> ftile_bytes(opt.flux_blob ? (size_t)ntiles : 0);
  out.ftile_nfaces.assign(opt.flux_blob ? (size_t)ntiles : 0, 0); out.ftile_nhalo = out.ftile_nfaces; out.ftile_maxdeg = out.ftile_nfaces;
  long long wf_min_total = 0, wf_est_total = 0, wf_v = 0, wf_n = 0, wf_steps = 0;
#pragma omp parallel reduction(+ : wf_min_total, wf_est_total, wf_v, wf_n, wf_steps)
  {
    std::vector<int> &lmap = lmap_t[omp_get_thread_num()];
    struct Ent { int tt, p1, p0, face; int nbr; /* tile-local: own i or n + halo k */ int fid; uint32_t sign; };
    std::vector<Ent> ents, tile_ents;            /* tile_ents[i*md + j] */
    std::vector<int> deg, slot_of, hpos_of;      /* fid -> slot, halo k -> position */
    std::vector<unsigned char> cntv, cntn;       /* [group][16] */
    std::vector<int> grp_off, grp_list;          /* per face / per halo point: distinct groups */
    std::vector<int> cls; std::vector<std::vector<int>> members; /* class of an object; objects of a class */
    std::vector<Ent> ftile_ents; std::vector<int> fdeg, fface_of, fhalo_of, fslot_of, fhpos_of; std::vector<unsigned char> fghost; /* pseudo-flux blob */
    double tB[6] = {0, 0, 0, 0, 0, 0}; double tb0 = 0;
    const bool tprof = prof && omp_get_thread_num() == 0;
#define TB(i) do { if (tprof) { const double t_ = wall(); tB[i] += t_ - tb0; tb0 = t_; } } while (0)
#pragma omp for schedule(dynamic, 16) nowait
    for (int k = 0; k < ntiles; k++) {
      if (tprof) tb0 = wall();
      const int n = out.tile_npts[k], nf = out.tile_nfaces[k], nh = tnh[k], md = tmaxdeg[k];
      const int nslots = out.tile_nslots[k], nhpos = out.tile_nhpos[k], n_even = opt.gather ? CFDP_HALO_BASE_G(n) : CFDP_HALO_BASE(n); /* first halo position */
      const uint32_t npad = (uint32_t)align_up((size_t)n, 32);
      const int *P = &tb.pts[pt_off[k]];
      unsigned char *bl = &out.blob[out.tile_blob[k]];
      double *nrm = (double *)bl;
      uint32_t *hrows = (uint32_t *)(bl + blob_halo_off((uint32_t)nslots));
      uint32_t *ell = (uint32_t *)(bl + blob_adj_off((uint32_t)nslots, (uint32_t)nhpos));
      int *fids = &out.tile_face_ids[(size_t)out.tile_face_off[k]];
      int *hpts = &out.tile_halo_pts[(size_t)out.tile_halo_off[k]];
      for (size_t i = 0; i < (size_t)md * npad; i++) ell[i] = CFDP_ADJ_PAD;
      for (int j = 0; j < nhpos; j++) hrows[j] = 0xFFFFFFFFu;
      for (int i = 0; i < n; i++) lmap[P[i]] = i;
      int hcount = 0;
      for (int i = 0; i < n; i++) {
        const int p = P[i];
        for (long long a = g.off[p]; a < g.off[(size_t)p + 1]; a++) {
          const int q = g.nbr[a];
          if (lmap[q] < 0) { lmap[q] = n + hcount; hpts[hcount] = q; hcount++; }
        }
      }
      ASSERT(hcount == nh);
      /* sorted entry lists */
      tile_ents.assign((size_t)n * md, Ent{0, 0, 0, -1, -1, -1, 0});
      deg.assign((size_t)n, 0);
      for (int i = 0; i < n; i++) {
        const int p = P[i];
        ents.clear();
        for (long long a = g.off[p]; a < g.off[(size_t)p + 1]; a++) {
          const int q = g.nbr[a], f = g.face[a];
          int fid = eslot[a];
          if (fid < 0) { /* p is p1 of an internal face: numbered from q's (p0) side */
            for (long long c = g.off[q]; c < g.off[(size_t)q + 1]; c++) if (g.face[c] == f && eslot[c] >= 0) { fid = eslot[c]; break; }
            ASSERT(fid >= 0);
          } else {
            fids[fid] = f;
          }
          const int p0 = g.sign[a] ? q : p, p1 = g.sign[a] ? p : q;
          const int h0 = p0 >= nown ? 3 : (is_send[p0] ? 2 : 1), h1 = p1 >= nown ? 3 : (is_send[p1] ? 2 : 1);
          Ent en;
          en.tt = ((h0 == 2 || h1 == 2) ? 0 : 3) + (h0 == 3 ? 0 : (h1 == 3 ? 1 : 2)); /* rangelist.c:567-608 with one thread */
          en.p1 = p1; en.p0 = p0; en.face = f; en.nbr = lmap[q]; en.fid = fid; en.sign = g.sign[a];
          ents.push_back(en);
        }
        std::sort(ents.begin(), ents.end(), [](const Ent &x, const Ent &y) {
          if (x.tt != y.tt) return x.tt < y.tt;
          if (x.p1 != y.p1) return x.p1 < y.p1;
          if (x.p0 != y.p0) return x.p0 < y.p0;
          return x.face < y.face;
        });
        deg[i] = (int)ents.size();
        for (size_t j = 0; j < ents.size(); j++) tile_ents[(size_t)i * md + j] = ents[j];
      }
      for (int i = 0; i < n; i++) lmap[P[i]] = -1;
      for (int j = 0; j < nh; j++) lmap[hpts[j]] = -1;
      TB(0);

      /* ---- placement.  A half-warp (16 lanes = 16 consecutive tile points) at step j of the face walk
       * reads one normal and one var row per lane with 8-byte loads; two lanes collide when they read different
       * words of the same bank pair.  Row r / slot s sits in bank-pair class r mod 16 / s mod 16 (row and slot
       * strides are 7 and 3 words, both odd).  Own rows are fixed by the tile order; halo rows and face slots
       * are free: greedily give each the class that adds the fewest collisions over the groups it is read in. */
      /* the placement, for an adjacency given as (E_[i*MD + j], D_[i]): nH halo points -> HP (NHP positions), nF faces -> SL (NSL slots) */
      auto place_all = [&](const std::vector<Ent> &E_, const std::vector<int> &D_, int MD, int nH, int nF,
                           std::vector<int> &HP, int NHP, std::vector<int> &SL, int NSL, bool rows64, int halo_base) {
      SL.assign((size_t)nF, -1); HP.assign((size_t)nH, -1);
      const int nhw = (n + 15) / 16, ngrp = nhw * MD;
      if (!place_by_bank) {
        for (int f = 0; f < nF; f++) SL[f] = f;
        for (int h = 0; h < nH; h++) HP[h] = h;
      } else {
        cntv.assign((size_t)ngrp * 16, 0); cntn.assign((size_t)ngrp * 16, 0);
        /* own rows: distinct rows per group */
        for (int hw = 0; hw < nhw; hw++)
          for (int j = 0; j < MD; j++) {
            unsigned char *cv = &cntv[((size_t)hw * MD + j) * 16];
            int seen[16], ns = 0;
            for (int l = 0; l < 16; l++) {
              const int i = hw * 16 + l;
              if (i >= n || j >= D_[i]) continue;
              const int r = E_[(size_t)i * MD + j].nbr;
              if (r >= n) continue;
              bool dup = false;
              for (int t = 0; t < ns; t++) if (seen[t] == r) dup = true;
              if (!dup) { seen[ns++] = r; cv[r & 15]++; }
            }
          }
        auto place = [&](int nobj, bool is_face, std::vector<unsigned char> &cnt, std::vector<int> &pos_of, int npos, int base_mod) {
          /* object -> distinct groups */
          grp_off.assign((size_t)nobj + 1, 0);
          auto obj_of = [&](const Ent &e) { return is_face ? e.fid : (e.nbr >= n ? e.nbr - n : -1); };
          for (int pass = 0; pass < 2; pass++) {
            if (pass == 1) { for (int o = 0; o < nobj; o++) grp_off[(size_t)o + 1] += grp_off[o]; grp_list.assign((size_t)grp_off[nobj], -1); }
            std::vector<int> fill;
            if (pass == 1) fill.assign(grp_off.begin(), grp_off.end() - 1);
            for (int i = 0; i < n; i++)
              for (int j = 0; j < D_[i]; j++) {
                const int o = obj_of(E_[(size_t)i * MD + j]);
                if (o < 0) continue;
                const int gidx = (i / 16) * MD + j;
                if (pass == 0) grp_off[(size_t)o + 1]++;
                else {
                  bool dup = false;
                  for (int t = grp_off[o]; t < fill[o]; t++) if (grp_list[t] == gidx) dup = true;
                  if (!dup) grp_list[fill[o]++] = gidx;
                }
              }
          }
          const int cap = npos / 16;
          int used[16] = {0};
          cls.assign((size_t)nobj, -1);
          std::vector<int> gmx;
          for (int o = 0; o < nobj; o++) {
            /* the fullest class of each group this object is read in: once per object, not once per candidate class */
            gmx.assign((size_t)(grp_off[(size_t)o + 1] - grp_off[o]), 0);
            for (int t = grp_off[o]; t < grp_off[(size_t)o + 1]; t++) {
              const int gi = grp_list[t];
              if (gi < 0) continue;
              const unsigned char *cc = &cnt[(size_t)gi * 16];
              int mx = 0;
              for (int x = 0; x < 16; x++) mx = std::max(mx, (int)cc[x]);
              gmx[(size_t)(t - grp_off[o])] = mx;
            }
            int best = -1, best_cost = 1 << 30;
            for (int c0 = 0; c0 < 16; c0++) {
              const int c = (c0 + o) & 15; /* rotate the tie-break so that classes fill evenly */
              if (used[c] >= cap) continue;
              int cost = 0;
              for (int t = grp_off[o]; t < grp_off[(size_t)o + 1]; t++) {
                const int gi = grp_list[t];
                if (gi < 0) continue;
                const int ccc = cnt[(size_t)gi * 16 + c];
                if (ccc + 1 > gmx[(size_t)(t - grp_off[o])]) cost += 4;          /* raises this group's wavefront count */
                cost += ccc;                            /* prefer emptier classes */
              }
              if (cost < best_cost) { best_cost = cost; best = c; }
            }
            ASSERT(best >= 0);
            for (int t = grp_off[o]; t < grp_off[(size_t)o + 1]; t++) if (grp_list[t] >= 0) cnt[(size_t)grp_list[t] * 16 + best]++;
            cls[o] = best;
            used[best]++;
          }
          /* refinement: an object that alone holds a group at its maximum moves to a class where it does not, into a
           * free position or by trading places with an object for which the trade costs nothing.  The objective is the
           * estimator's: the sum over groups of the fullest class. */
          auto gmax = [&](int gi) { const unsigned char *cc = &cnt[(size_t)gi * 16]; int mx = 0; for (int x = 0; x < 16; x++) mx = std::max(mx, (int)cc[x]); return mx; };
          auto n_at = [&](int gi, int v) { const unsigned char *cc = &cnt[(size_t)gi * 16]; int k = 0; for (int x = 0; x < 16; x++) k += cc[x] == v; return k; };
          auto move_delta = [&](int o, int from, int to) { /* change of the objective when o goes from class `from` to `to` (counts include o in `from`) */
            int d = 0;
            for (int t = grp_off[o]; t < grp_off[(size_t)o + 1]; t++) {
              const int gi = grp_list[t];
              if (gi < 0) continue;
              const unsigned char *cc = &cnt[(size_t)gi * 16];
              const int mx = gmax(gi);
              const int mx_wo = (cc[from] == mx && n_at(gi, mx) == 1) ? mx - 1 : mx; /* maximum without o */
              d += std::max(mx_wo, cc[to] + 1) - mx;
            }
            return d;
          };
          auto apply_move = [&](int o, int from, int to) {
            for (int t = grp_off[o]; t < grp_off[(size_t)o + 1]; t++) {
              const int gi = grp_list[t];
              if (gi < 0) continue;
              cnt[(size_t)gi * 16 + from]--; cnt[(size_t)gi * 16 + to]++;
            }
            cls[o] = to;
          };
          for (int round = 0; round < opt.refine_rounds; round++) {
            members.assign(16, std::vector<int>());
            for (int o = 0; o < nobj; o++) members[(size_t)cls[o]].push_back(o);
            int improved = 0;
            for (int o = 0; o < nobj; o++) {
              const int c = cls[o];
              bool critical = false;
              for (int t = grp_off[o]; t < grp_off[(size_t)o + 1] && !critical; t++) {
                const int gi = grp_list[t];
                if (gi < 0) continue;
                const int mx = gmax(gi);
                critical = mx > 1 && cnt[(size_t)gi * 16 + c] == mx && n_at(gi, mx) == 1;
              }
              if (!critical) continue;
              int best_d = 0, best_c = -1, best_partner = -1;
              for (int c2 = 0; c2 < 16; c2++) {
                if (c2 == c) continue;
                const int d1 = move_delta(o, c, c2);
                if (d1 >= 0) continue;
                if (used[c2] < cap) { if (d1 < best_d) { best_d = d1; best_c = c2; best_partner = -1; } continue; }
                /* trade: apply o's move, then look for a partner in c2 whose move to c keeps the total below zero */
                apply_move(o, c, c2);
                int tries = 0;
                for (int q : members[(size_t)c2]) {
                  if (cls[q] != c2 || q == o) continue;
                  if (++tries > 24) break;
                  const int d2 = move_delta(q, c2, c);
                  if (d1 + d2 < best_d) { best_d = d1 + d2; best_c = c2; best_partner = q; }
                }
                apply_move(o, c2, c);
              }
              if (best_c < 0) continue;
              apply_move(o, c, best_c);
              if (best_partner >= 0) { apply_move(best_partner, best_c, c); members[(size_t)c].push_back(best_partner); }
              else { used[c]--; used[best_c]++; }
              members[(size_t)best_c].push_back(o);
              improved++;
            }
            if (!improved) break;
          }
          int fillc[16] = {0};
          for (int o = 0; o < nobj; o++) {
            /* position with (base_mod + pos) mod 16 == class */
            const int first = ((cls[o] - base_mod) % 16 + 16) % 16;
            pos_of[o] = first + 16 * fillc[cls[o]]++;
          }
        };
        /* gather mode: the face walk reads a var row as 16-byte pieces (ld.shared.v2.f64); a quarter warp (8 lanes = 8
         * consecutive tile points) is conflict free when its rows fall into 8 different 16-byte bank groups, and with the
         * TMA 64-byte swizzle the group of a piece is a function of (position mod 8) only: class = position mod 8 */
        auto place_halo8 = [&](std::vector<int> &pos_of, int npos) {
          const int nq = (n + 7) / 8, ng = nq * MD;
          std::vector<unsigned char> c8((size_t)ng * 8, 0);
          for (int q = 0; q < nq; q++)
            for (int j = 0; j < MD; j++) {
              unsigned char *cv = &c8[((size_t)q * MD + j) * 8];
              int seen[8], ns = 0;
              for (int l = 0; l < 8; l++) {
                const int i = q * 8 + l;
                if (i >= n || j >= D_[i]) continue;
                const int r = E_[(size_t)i * MD + j].nbr;
                if (r >= n) continue;
                bool dup = false;
                for (int t = 0; t < ns; t++) if (seen[t] == r) dup = true;
                if (!dup) { seen[ns++] = r; cv[r & 7]++; }
              }
            }
          grp_off.assign((size_t)nH + 1, 0);
          std::vector<int> fill;
          for (int pass = 0; pass < 2; pass++) {
            if (pass == 1) { for (int o = 0; o < nH; o++) grp_off[(size_t)o + 1] += grp_off[o]; grp_list.assign((size_t)grp_off[nH], -1); fill.assign(grp_off.begin(), grp_off.end() - 1); }
            for (int i = 0; i < n; i++)
              for (int j = 0; j < D_[i]; j++) {
                const int r = E_[(size_t)i * MD + j].nbr;
                if (r < n) continue;
                const int o = r - n, gidx = (i / 8) * MD + j;
                if (pass == 0) grp_off[(size_t)o + 1]++;
                else {
                  bool dup = false;
                  for (int t = grp_off[o]; t < fill[o]; t++) if (grp_list[t] == gidx) dup = true;
                  if (!dup) grp_list[fill[o]++] = gidx;
                }
              }
          }
          const int cap = npos / 8;
          int used[8] = {0}, fillc[8] = {0};
          for (int o = 0; o < nH; o++) {
            int best = -1, best_cost = 1 << 30;
            for (int c0 = 0; c0 < 8; c0++) {
              const int c = (c0 + o) & 7;
              if (used[c] >= cap) continue;
              int cost = 0;
              for (int t = grp_off[o]; t < grp_off[(size_t)o + 1]; t++) {
                const int gi = grp_list[t];
                if (gi < 0) continue;
                const unsigned char *cc = &c8[(size_t)gi * 8];
                int mx = 0;
                for (int x = 0; x < 8; x++) mx = std::max(mx, (int)cc[x]);
                if (cc[c] + 1 > mx) cost += 4;
                cost += cc[c];
              }
              if (cost < best_cost) { best_cost = cost; best = c; }
            }
            ASSERT(best >= 0);
            for (int t = grp_off[o]; t < grp_off[(size_t)o + 1]; t++) if (grp_list[t] >= 0) c8[(size_t)grp_list[t] * 8 + best]++;
            used[best]++;
            pos_of[o] = best + 8 * fillc[best]++;   /* the first halo position is a multiple of 16 */
          }
        };
        if (rows64) place_halo8(HP, NHP); else place(nH, false, cntv, HP, NHP, halo_base & 15);
        place(nF, true, cntn, SL, NSL, 0);
      }
      };
      place_all(tile_ents, deg, md, nh, nf, hpos_of, nhpos, slot_of, nslots, opt.gather != 0, n_even);
      TB(1);

      /* ---- emit */
      {
        std::vector<unsigned char> used_slot((size_t)nslots, 0);
        for (int f = 0; f < nf; f++) used_slot[(size_t)slot_of[f]] = 1;
        int z = nslots - 1;
        while (z >= 0 && used_slot[(size_t)z]) z--;
        ASSERT(z >= 0);
        out.tile_zslot[k] = z;
      }
      for (int f = 0; f < nf; f++) {
        const int sl = slot_of[f], gf = fids[f];
        ASSERT(sl >= 0 && sl < nslots);
        nrm[3 * sl + 0] = sd->fnormal[gf][0]; nrm[3 * sl + 1] = sd->fnormal[gf][1]; nrm[3 * sl + 2] = sd->fnormal[gf][2];
      }
      for (int h = 0; h < nh; h++) { ASSERT(hpos_of[h] >= 0 && hpos_of[h] < nhpos); hrows[hpos_of[h]] = (uint32_t)out.row_of_point[hpts[h]]; }
      for (int i = 0; i < n; i++)
        for (int j = 0; j < deg[i]; j++) {
          const Ent &e = tile_ents[(size_t)i * md + j];
          const uint32_t loc = e.nbr < n ? (uint32_t)e.nbr : (uint32_t)(n_even + hpos_of[e.nbr - n]);
          const uint32_t ghost = (e.nbr >= n && hpts[e.nbr - n] >= sd->nownpoints) ? 0x8000u : 0u;
          ASSERT(loc < 0x8000u);
          ell[(size_t)j * npad + i] = loc | ghost | ((uint32_t)slot_of[e.fid] << 16) | (e.sign << 31);
        }

      TB(2);
      /* ---- the pseudo-flux blob of the tile (flux.c:179-190 with one thread): an own point receives -flux from the
       * faces where it is p1 and +flux from the faces where it is p0 and p1 is a ghost; nothing else is stored.
       * Face slots and halo positions are placed by bank for this adjacency like those of the gradient blob. */
      if (opt.flux_blob) {
        auto contributes = [&](const Ent &e) { return e.sign != 0 || (e.nbr >= n && hpts[e.nbr - n] >= nown); };
        /* compact numbering of the faces and halo points the contributing entries reference */
        fface_of.assign((size_t)nf, -1); fhalo_of.assign((size_t)nh, -1);
        fdeg.assign((size_t)n, 0);
        int fmd = 0, nff = 0, nfh = 0;
        for (int i = 0; i < n; i++) {
          int c = 0;
          for (int j = 0; j < deg[i]; j++) c += contributes(tile_ents[(size_t)i * md + j]) ? 1 : 0;
          fdeg[i] = c; fmd = std::max(fmd, c);
        }
        ftile_ents.assign((size_t)n * std::max(fmd, 1), Ent{0, 0, 0, -1, -1, -1, 0});
        fghost.assign((size_t)n * std::max(fmd, 1), 0);
        for (int i = 0; i < n; i++) {
          int c = 0;
          for (int j = 0; j < deg[i]; j++) {
            const Ent &e = tile_ents[(size_t)i * md + j];
            if (!contributes(e)) continue;
            Ent fe = e;
            if (fface_of[e.fid] < 0) fface_of[e.fid] = nff++;
            fe.fid = fface_of[e.fid];
            if (e.nbr >= n) {
              if (fhalo_of[e.nbr - n] < 0) fhalo_of[e.nbr - n] = nfh++;
              fe.nbr = n + fhalo_of[e.nbr - n];
              fghost[(size_t)i * fmd + c] = hpts[e.nbr - n] >= nown ? 1 : 0;
            }
            ftile_ents[(size_t)i * fmd + c] = fe;
            c++;
          }
        }
        const int nfslots = (int)align_up((size_t)nff, 16), nfhpos = (int)align_up((size_t)nfh, 16);
        const int fbase = CFDP_HALO_BASE(n);   /* the pseudo-flux kernel keeps 72-byte rows: its own halo base and 16 classes */
        place_all(ftile_ents, fdeg, std::max(fmd, 1), nfh, nff, fhpos_of, nfhpos, fslot_of, nfslots, false, fbase);
        std::vector<unsigned char> &fb = ftile_bytes[k];
        fb.assign(blob_size((uint32_t)nfslots, (uint32_t)nfhpos, (uint32_t)fmd, npad), 0);
        double *fn = (double *)fb.data();
        uint32_t *fh = (uint32_t *)(fb.data() + blob_halo_off((uint32_t)nfslots));
        uint32_t *fe = (uint32_t *)(fb.data() + blob_adj_off((uint32_t)nfslots, (uint32_t)nfhpos));
        for (size_t i = 0; i < (size_t)fmd * npad; i++) fe[i] = CFDP_ADJ_PAD;
        for (int j = 0; j < nfhpos; j++) fh[j] = 0xFFFFFFFFu;
        for (int f = 0; f < nf; f++)
          if (fface_of[f] >= 0) {
            const int gf = fids[f], sl = fslot_of[fface_of[f]];
            ASSERT(sl >= 0 && sl < nfslots);
            fn[3 * sl] = sd->fnormal[gf][0]; fn[3 * sl + 1] = sd->fnormal[gf][1]; fn[3 * sl + 2] = sd->fnormal[gf][2];
          }
        for (int h = 0; h < nh; h++)
          if (fhalo_of[h] >= 0) { ASSERT(fhpos_of[fhalo_of[h]] >= 0 && fhpos_of[fhalo_of[h]] < nfhpos); fh[fhpos_of[fhalo_of[h]]] = (uint32_t)out.row_of_point[hpts[h]]; }
        for (int i = 0; i < n; i++)
          for (int c = 0; c < fdeg[i]; c++) {
            const Ent &e = ftile_ents[(size_t)i * fmd + c];
            const uint32_t loc = e.nbr < n ? (uint32_t)e.nbr : (uint32_t)(fbase + fhpos_of[e.nbr - n]);
            fe[(size_t)c * npad + i] = loc | (fghost[(size_t)i * fmd + c] ? 0x8000u : 0u) | ((uint32_t)fslot_of[e.fid] << 16) | (e.sign << 31);
          }
        out.ftile_nfaces[k] = nfslots; out.ftile_nhalo[k] = nfhpos; out.ftile_maxdeg[k] = fmd;
      }

      TB(3);
      /* shared-memory wavefront estimate of the face walk: 7 var words + 3 normal words per face end */
      for (int w0 = 0; w0 < n; w0 += 16)
        for (int j = 0; j < md; j++) {
          int cnt_v[16] = {0}, cnt_n[16] = {0}, nact = 0;
          uint32_t seen_v[16], seen_n[16]; int nsv = 0, nsn = 0;
          for (int l = 0; l < 16; l++) {
            const int i = w0 + l;
            if (i >= n) break;
            const uint32_t e = ell[(size_t)j * npad + i];
            if (e == CFDP_ADJ_PAD) continue;
            nact++;
            const uint32_t r = e & 0x7FFFu, sl = (e >> 16) & 0x7FFFu;
            bool dup = false;
            for (int t = 0; t < nsv; t++) if (seen_v[t] == r) dup = true;
            if (!dup) { seen_v[nsv++] = r; cnt_v[r & 15]++; }
            dup = false;
            for (int t = 0; t < nsn; t++) if (seen_n[t] == sl) dup = true;
            if (!dup) { seen_n[nsn++] = sl; cnt_n[sl & 15]++; }
          }
          if (!nact) continue;
          int mv = 0, mn = 0;
          for (int c = 0; c < 16; c++) { mv = std::max(mv, cnt_v[c]); mn = std::max(mn, cnt_n[c]); }
          wf_min_total += 10; wf_est_total += 7 * mv + 3 * mn; wf_v += mv; wf_n += mn; wf_steps++;
        }
      TB(4);
    }
    if (tprof) fprintf(stderr, "  pass B thread 0: entries+sort %.2f  place %.2f  emit %.2f  flux blob %.2f  estimate %.2f s\n", tB[0], tB[1], tB[2], tB[3], tB[4]);
#undef TB
  }
  lap("pass B (placement + emit)");
  out.lds_wavefronts_min = wf_min_total; out.lds_wavefronts_est = wf_est_total;
  if (prof && wf_steps) fprintf(stderr, "  schedule: half-warp steps %lld, mean depth of the worst bank-pair class: var rows %.3f, normals %.3f\n", wf_steps, (double)wf_v / wf_steps, (double)wf_n / wf_steps);
  if (opt.flux_blob) {
    out.ftile_blob.assign((size_t)ntiles + 1, 0);
    for (int k = 0; k < ntiles; k++) out.ftile_blob[(size_t)k + 1] = out.ftile_blob[k] + ftile_bytes[k].size();
    out.fblob.resize((size_t)out.ftile_blob[ntiles]);
#pragma omp parallel for schedule(static)
    for (int k = 0; k < ntiles; k++) {
      if (!ftile_bytes[k].empty()) memcpy(&out.fblob[out.ftile_blob[k]], ftile_bytes[k].data(), ftile_bytes[k].size());
      std::vector<unsigned char>().swap(ftile_bytes[k]);
    }
  }
}
