// grad_sched.cuh -- work schedule of the single-recompute backward (grad_both_kernel).
//
// One sweep over the (row block i, column tile j) grid of S forms every G tile ONCE and feeds it to
// two accumulations with opposite residency:
//     dI_i += G(i,j)   T_j     wants i fixed while j sweeps   (row-resident accumulator)
//     dT_j += G(i,j)^T I_i     wants j fixed while i sweeps   (column-resident accumulator)
// Three roles share the SMs: NP "producer" slots (S + softmax -> G, each paired with a dI consumer
// that follows it tile by tile) and NQ >= NP "dT consumers" fed through an L2-resident ring.
//
// The schedule is a skewed (systolic) sweep in nominal time steps.  Work is cut in PHASES; a phase
// takes `n_rows` row blocks, splits every row's column range in `n_seg` segments of `cs` tiles
// (the last may be shorter) and deals the n_rows * n_seg "virtual rows" v = seg * n_rows + r to the
// producer slots in WAVES of `np` (v = wave * np + slot).  A wave lasts `cs` steps; at step u
// slot a works on column  seg*cs + (a + u) mod cs  of its virtual row.  Hence at every step the
// np slots hit np DIFFERENT columns, and a column receives its tiles of a wave from consecutive
// slots at consecutive steps: slot a at step u with a + u = e ("run" e of the wave, 0 <= e <
// cs + np - 1; runs e and e + cs hit the same column index).  Run ids are numbered globally and dealt
// round-robin to the dT consumers: runs that overlap in time have distinct consumers because at
// most np <= NQ consecutive run ids are active at any step.  Inside a run the tiles of one column
// segment are contiguous: a PIECE (wave, column, wrapped?) -- the unit a dT consumer accumulates
// in TMEM before it adds the block to the column's partial sum in global memory.  The pieces of
// a column are ranked by (wave, wrapped): rank k adds after rank k - 1 (fixed order => dT is
// bit-reproducible), the last rank writes the final rows.
//
// Every wait in the kernel points to work with a smaller nominal time (producers emit and
// consumers accept tiles in non-decreasing time), so the ring protocol cannot deadlock.
// Everything here is closed form and shared by host (plan, CPU tests) and device.
#pragma once
#include <cstdint>

#ifndef __CUDACC__
#define VLP_SCHED_HD inline
#else
#define VLP_SCHED_HD __host__ __device__ inline
#endif

namespace vlp {

constexpr int SCHED_MAX_PHASES = 2;
constexpr int SCHED_MAX_SEG = 6;

struct SchedPhase {
  int t0;        // first nominal step
  int n_waves;
  int cs;        // steps per wave = tiles per column segment
  int n_seg;     // column segments per row block
  int row0;      // first row block
  int n_rows;    // row blocks
  int np;        // producer slots in use (<= cs)
  int part0;     // first dI partial slot (n_seg > 1), else -1
  int run0;      // id of the phase's first run
  int wave0;     // global index of the phase's first wave
};

struct Sched {
  SchedPhase ph[SCHED_MAX_PHASES];
  int n_ph;
  int R, C;        // row blocks, column tiles
  int np, nq;      // producer slots, dT consumers
  int t_total;     // nominal steps
  int n_parts;     // dI partial slots (row blocks whose column range is split)
  int n_runs;
  int n_waves;
};

VLP_SCHED_HD int sched_min(int a, int b) { return a < b ? a : b; }
VLP_SCHED_HD int sched_max(int a, int b) { return a > b ? a : b; }

// ---- host: build the schedule ----------------------------------------------------------------
inline Sched make_sched(int R, int C, int np, int nq) {
  Sched s = {};
  s.R = R;
  s.C = C;
  s.np = np;
  s.nq = nq;
  int t = 0, run = 0, wave = 0, parts = 0, row = 0;
  auto add = [&](int n_rows, int n_seg, int cs, int npp) {
    SchedPhase& p = s.ph[s.n_ph++];
    p.t0 = t;
    p.cs = cs;
    p.n_seg = n_seg;
    p.row0 = row;
    p.n_rows = n_rows;
    p.np = npp;
    p.n_waves = (n_rows * n_seg + npp - 1) / npp;
    p.part0 = n_seg > 1 ? parts : -1;
    p.run0 = run;
    p.wave0 = wave;
    if (n_seg > 1) parts += n_rows * n_seg;
    t += p.n_waves * cs;
    run += p.n_waves * (cs + npp - 1);
    wave += p.n_waves;
    row += n_rows;
  };
  // phase A: whole rows, as many full waves as there are
  const int np_a = sched_min(np, C);
  const int waves_a = R / np_a;
  if (waves_a > 0) add(waves_a * np_a, 1, C, np_a);
  // phase B: the remaining rows, column range split so that the slots stay busy
  const int rb = R - row;
  if (rb > 0) {
    int best_seg = 1, best_t = 1 << 30;
    for (int sc = 1; sc <= SCHED_MAX_SEG && sc <= C; ++sc) {
      const int cs = (C + sc - 1) / sc;
      if ((sc - 1) * cs >= C) continue;   // an empty last segment
      const int npp = sched_min(np, cs);
      const int waves = (rb * sc + npp - 1) / npp;
      const int tt = waves * cs + 4 * (sc - 1);   // small bias towards fewer partial blocks
      if (tt < best_t) {
        best_t = tt;
        best_seg = sc;
      }
    }
    const int cs = (C + best_seg - 1) / best_seg;
    add(rb, best_seg, cs, sched_min(np, cs));
  }
  s.t_total = t;
  s.n_runs = run;
  s.n_waves = wave;
  s.n_parts = parts;
  return s;
}

// ---- producer side: the virtual row of slot `a` in wave `w` of phase `p` ---------------------------
struct VRow {
  int rb;        // row block
  int seg;
  int c_lo;      // first column tile of the segment
  int c_n;       // tiles in the segment
  int part;      // dI partial slot or -1 (whole row: final rows written by the dI consumer)
};
VLP_SCHED_HD bool sched_vrow(const Sched& s, const SchedPhase& p, int w, int a, VRow& vr) {
  if (a >= p.np) return false;
  const int v = w * p.np + a;
  if (v >= p.n_rows * p.n_seg) return false;
  vr.seg = v / p.n_rows;
  vr.rb = p.row0 + (v - vr.seg * p.n_rows);
  vr.c_lo = vr.seg * p.cs;
  vr.c_n = sched_min(p.cs, s.C - vr.c_lo);
  vr.part = p.part0 >= 0 ? p.part0 + v : -1;
  return true;
}
// column tile of slot `a` at step u of a wave (or -1: idle step of a short last segment)
VLP_SCHED_HD int sched_col(const SchedPhase& p, const VRow& vr, int a, int u) {
  int cidx = a + u;
  if (cidx >= p.cs) cidx -= p.cs;
  return cidx < vr.c_n ? vr.c_lo + cidx : -1;
}

// ---- dT consumer side: pieces -------------------------------------------------------------------
struct Piece {
  int col;           // column tile
  int gw;            // global wave index
  int wrapped;       // second run of this column index in the wave
  int a_hi, a_lo;    // producer slots, visited from a_hi down to a_lo
  int t_hi;          // nominal step of slot a_hi's tile (slot a: t_hi + (a_hi - a))
  int rb_hi;         // row block of slot a_hi's tile (slot a: rb_hi - (a_hi - a))
};

// slots of segment `seg` in wave w of phase p, intersected with run e: false when empty
VLP_SCHED_HD bool sched_piece_slots(const Sched& s, const SchedPhase& p, int w, int e, int seg,
                                    int& a_lo, int& a_hi) {
  const int V = p.n_rows * p.n_seg;
  const int v0 = w * p.np;
  int lo = sched_max(0, e - p.cs + 1);
  int hi = sched_min(p.np - 1, e);
  hi = sched_min(hi, V - 1 - v0);
  lo = sched_max(lo, seg * p.n_rows - v0);
  hi = sched_min(hi, (seg + 1) * p.n_rows - 1 - v0);
  if (lo > hi) return false;
  int cidx = e;
  if (cidx >= p.cs) cidx -= p.cs;
  const int c_lo = seg * p.cs;
  if (c_lo >= s.C || cidx >= sched_min(p.cs, s.C - c_lo)) return false;
  a_lo = lo;
  a_hi = hi;
  return true;
}

// does column tile `col` receive a piece in wave w of phase p (first / wrapped run)?
VLP_SCHED_HD bool sched_piece_exists(const Sched& s, const SchedPhase& p, int w, int col, int wrapped) {
  const int seg = col / p.cs;
  if (seg >= p.n_seg) return false;
  const int e = col - seg * p.cs + (wrapped ? p.cs : 0);
  if (e > p.cs + p.np - 2) return false;
  int a_lo, a_hi;
  return sched_piece_slots(s, p, w, e, seg, a_lo, a_hi);
}

// rank of piece (gw, wrapped) among the pieces of its column, and their number
VLP_SCHED_HD void sched_piece_rank(const Sched& s, int col, int gw, int wrapped, int& rank, int& total) {
  rank = 0;
  total = 0;
  for (int pi = 0; pi < s.n_ph; ++pi) {
    const SchedPhase& p = s.ph[pi];
    for (int w = 0; w < p.n_waves; ++w)
      for (int wr = 0; wr < 2; ++wr)
        if (sched_piece_exists(s, p, w, col, wr)) {
          const int g = p.wave0 + w;
          if (g < gw || (g == gw && wr < wrapped)) ++rank;
          ++total;
        }
  }
}

// iterates the pieces of dT consumer q in the order it works through them
struct PieceIter {
  const Sched* s;
  int q;
  int pi, w, e, seg;   // phase, wave in phase, run in wave, next segment to try (descending)
  bool started;
  VLP_SCHED_HD PieceIter(const Sched& sched, int consumer) : s(&sched), q(consumer), pi(0), w(0), e(-1), seg(-1), started(false) {}
  VLP_SCHED_HD static int first_run(const SchedPhase& p, int w, int q, int nq) {
    const int base = p.run0 + w * (p.cs + p.np - 1);
    int r = (q - base) % nq;
    if (r < 0) r += nq;
    return r;
  }
  VLP_SCHED_HD bool next(Piece& pc) {
    while (pi < s->n_ph) {
      const SchedPhase& p = s->ph[pi];
      if (w >= p.n_waves) {
        ++pi;
        w = 0;
        e = -1;
        continue;
      }
      const int n_run = p.cs + p.np - 1;
      if (e < 0) {
        e = first_run(p, w, q, s->nq);
        seg = p.n_seg - 1;
      }
      if (e >= n_run) {
        ++w;
        e = -1;
        continue;
      }
      while (seg >= 0) {
        const int sg = seg--;
        int a_lo, a_hi;
        if (sched_piece_slots(*s, p, w, e, sg, a_lo, a_hi)) {
          int cidx = e;
          if (cidx >= p.cs) cidx -= p.cs;
          pc.col = sg * p.cs + cidx;
          pc.gw = p.wave0 + w;
          pc.wrapped = e >= p.cs ? 1 : 0;
          pc.a_hi = a_hi;
          pc.a_lo = a_lo;
          pc.t_hi = p.t0 + w * p.cs + (e - a_hi);
          pc.rb_hi = p.row0 + (w * p.np + a_hi - sg * p.n_rows);
          return true;
        }
      }
      e += s->nq;
      seg = p.n_seg - 1;
    }
    return false;
  }
};

}  // namespace vlp
