// Brute-force check of the single-recompute backward's schedule (csrc/grad_sched.cuh), host only.
// Built and run by tests/test_grad_sched.py (g++; no CUDA needed).  Properties checked per shape:
//   1. producers: every (row block, column tile) exactly once; one tile per slot and step; a slot's
//      tiles of a wave share the row block; at every step the slots hit distinct columns;
//   2. dT consumers: their pieces cover exactly the producers' tiles (same slot, step, row, column);
//      every consumer accepts tiles in strictly increasing nominal time;
//   3. pieces of a column: ranks are 0..total-1, unique, ordered by nominal time;
//   4. the ring protocol (depth `ring` tiles per producer slot, consumers in piece order, producers
//      in time order, ordered column flushes) completes: event simulation, no deadlock.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <set>
#include <tuple>
#include <vector>
#include "grad_sched.cuh"

using namespace vlp;

struct Tile {
  int a, t, rb, col;
};

static int check(int R, int C, int np, int nq, int ring) {
  Sched s = make_sched(R, C, np, nq);
  // ---- producers ----
  std::map<std::pair<int, int>, int> seen;            // (rb, col) -> count
  std::map<std::pair<int, int>, Tile> by_slot_time;   // (a, t)
  std::map<int, std::set<int>> cols_at;               // t -> columns
  long long n_tiles = 0;
  for (int pi = 0; pi < s.n_ph; ++pi) {
    const SchedPhase& p = s.ph[pi];
    if (p.np > p.cs || p.np > nq) {
      printf("FAIL %dx%d: np %d > cs %d or nq %d\n", R, C, p.np, p.cs, nq);
      return 1;
    }
    for (int a = 0; a < s.np; ++a)
      for (int w = 0; w < p.n_waves; ++w) {
        VRow vr;
        if (!sched_vrow(s, p, w, a, vr)) continue;
        if (vr.rb < 0 || vr.rb >= R || vr.c_n <= 0) {
          printf("FAIL %dx%d: bad virtual row\n", R, C);
          return 1;
        }
        for (int u = 0; u < p.cs; ++u) {
          const int col = sched_col(p, vr, a, u);
          if (col < 0) continue;
          const int t = p.t0 + w * p.cs + u;
          if (col >= C || t >= s.t_total) {
            printf("FAIL %dx%d: col/t out of range\n", R, C);
            return 1;
          }
          ++seen[{vr.rb, col}];
          if (by_slot_time.count({a, t})) {
            printf("FAIL %dx%d: slot %d has two tiles at t=%d\n", R, C, a, t);
            return 1;
          }
          by_slot_time[{a, t}] = {a, t, vr.rb, col};
          if (!cols_at[t].insert(col).second) {
            printf("FAIL %dx%d: two slots on column %d at t=%d\n", R, C, col, t);
            return 1;
          }
          ++n_tiles;
        }
      }
  }
  if (n_tiles != (long long)R * C || (long long)seen.size() != (long long)R * C) {
    printf("FAIL %dx%d: %lld tiles, %zu distinct (want %lld)\n", R, C, n_tiles, seen.size(),
           (long long)R * C);
    return 1;
  }
  // ---- consumers ----
  std::set<std::pair<int, int>> consumed;
  std::map<int, std::vector<std::tuple<int, int, int>>> col_pieces;   // col -> (rank, total, first t)
  std::vector<std::vector<Piece>> pieces(nq);
  for (int q = 0; q < nq; ++q) {
    PieceIter it(s, q);
    Piece pc;
    int last_t = -1;
    while (it.next(pc)) {
      pieces[q].push_back(pc);
      int rank, total;
      sched_piece_rank(s, pc.col, pc.gw, pc.wrapped, rank, total);
      col_pieces[pc.col].push_back({rank, total, pc.t_hi});
      for (int a = pc.a_hi; a >= pc.a_lo; --a) {
        const int t = pc.t_hi + (pc.a_hi - a), rb = pc.rb_hi - (pc.a_hi - a);
        auto f = by_slot_time.find({a, t});
        if (f == by_slot_time.end() || f->second.rb != rb || f->second.col != pc.col) {
          printf("FAIL %dx%d: consumer %d tile (a=%d,t=%d,rb=%d,col=%d) not produced\n", R, C, q, a,
                 t, rb, pc.col);
          return 1;
        }
        if (!consumed.insert({a, t}).second) {
          printf("FAIL %dx%d: tile consumed twice\n", R, C);
          return 1;
        }
        if (t <= last_t) {
          printf("FAIL %dx%d: consumer %d time not increasing (%d after %d)\n", R, C, q, t, last_t);
          return 1;
        }
        last_t = t;
      }
    }
  }
  if ((long long)consumed.size() != n_tiles) {
    printf("FAIL %dx%d: consumed %zu of %lld tiles\n", R, C, consumed.size(), n_tiles);
    return 1;
  }
  for (auto& kv : col_pieces) {
    auto v = kv.second;
    std::sort(v.begin(), v.end());
    for (size_t k = 0; k < v.size(); ++k) {
      if (std::get<0>(v[k]) != (int)k || std::get<1>(v[k]) != (int)v.size()) {
        printf("FAIL %dx%d: column %d piece ranks inconsistent\n", R, C, kv.first);
        return 1;
      }
      if (k > 0 && std::get<2>(v[k]) <= std::get<2>(v[k - 1])) {
        printf("FAIL %dx%d: column %d piece ranks not in time order\n", R, C, kv.first);
        return 1;
      }
    }
  }
  if ((int)col_pieces.size() != C) {
    printf("FAIL %dx%d: %zu columns have pieces\n", R, C, col_pieces.size());
    return 1;
  }
  // ---- protocol simulation: each agent advances when its next action is enabled ----
  // producer a: emits its tiles in time order; tile k needs ring slot (t % ring) free, i.e. the
  // previous tile of that slot consumed.  consumer q: takes tiles in piece order; a piece's flush
  // needs the previous rank of its column flushed.
  std::vector<std::vector<Tile>> ptiles(s.np);
  for (auto& kv : by_slot_time) ptiles[kv.second.a].push_back(kv.second);
  std::vector<size_t> ppos(s.np, 0);
  std::set<std::pair<int, int>> in_ring, done;        // (a, t)
  std::vector<std::vector<int>> slot_last(s.np, std::vector<int>(ring, -1));
  std::map<int, int> col_turn;
  std::vector<size_t> cpiece(nq, 0);
  const int kStart = -1000000;
  std::vector<int> cslot(nq, kStart);                  // next slot within the piece (kStart: not begun)
  bool progress = true;
  long long total_done = 0;
  while (progress) {
    progress = false;
    for (int a = 0; a < s.np; ++a) {
      while (ppos[a] < ptiles[a].size()) {
        const Tile& tl = ptiles[a][ppos[a]];
        const int rs = tl.t % ring;
        const int prev = slot_last[a][rs];
        if (prev >= 0 && !done.count({a, prev})) break;
        slot_last[a][rs] = tl.t;
        in_ring.insert({a, tl.t});
        ++ppos[a];
        progress = true;
      }
    }
    for (int q = 0; q < nq; ++q) {
      while (cpiece[q] < pieces[q].size()) {
        const Piece& pc = pieces[q][cpiece[q]];
        if (cslot[q] == kStart) cslot[q] = pc.a_hi;
        bool blocked = false;
        while (cslot[q] >= pc.a_lo) {
          const int a = cslot[q], t = pc.t_hi + (pc.a_hi - a);
          if (!in_ring.count({a, t})) {
            blocked = true;
            break;
          }
          done.insert({a, t});
          ++total_done;
          --cslot[q];
          progress = true;
        }
        if (blocked) break;
        int rank, total;
        sched_piece_rank(s, pc.col, pc.gw, pc.wrapped, rank, total);
        if (col_turn[pc.col] != rank) break;   // predecessor not flushed yet
        col_turn[pc.col] = rank + 1;
        ++cpiece[q];
        cslot[q] = kStart;
        progress = true;
      }
    }
  }
  if (total_done != n_tiles) {
    printf("FAIL %dx%d (np %d nq %d ring %d): protocol stalls after %lld of %lld tiles\n", R, C, np,
           nq, ring, total_done, n_tiles);
    return 1;
  }
  const double ideal = (double)R * C / np;
  printf("ok %4d x %4d np %2d nq %2d: %d phases, %5d steps (ideal %.1f, %.1f%%), %d dI partials\n", R,
         C, np, nq, s.n_ph, s.t_total, ideal, 100.0 * ideal / s.t_total, s.n_parts);
  return 0;
}

int main(int argc, char** argv) {
  int bad = 0;
  if (argc >= 5) return check(atoi(argv[1]), atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), argc > 5 ? atoi(argv[5]) : 8);
  const int shapes[][2] = {{1, 1},   {1, 2},    {2, 1},    {3, 3},    {8, 8},    {2, 8},   {4, 32},
                           {8, 64},  {32, 32},  {49, 49},  {50, 50},  {64, 64},  {100, 7}, {7, 100},
                           {32, 256}, {64, 256}, {128, 256}, {256, 256}, {512, 512}, {63, 250}, {130, 131},
                           {48, 48}, {49, 256}, {51, 256}, {98, 256}, {16, 128}, {1, 300}, {300, 1}};
  for (auto& sh : shapes) {
    bad += check(sh[0], sh[1], 49, 50, 8);
    bad += check(sh[0], sh[1], 48, 48, 4);
  }
  bad += check(256, 256, 24, 26, 8);
  bad += check(37, 91, 5, 7, 2);
  bad += check(37, 91, 1, 1, 1);
  printf(bad ? "SCHED CHECK FAILED\n" : "SCHED CHECK PASSED\n");
  return bad ? 1 : 0;
}
