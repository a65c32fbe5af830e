#include "jsso_symbolic.h"

#include <algorithm>
#include <numeric>

namespace jsso {

namespace {
struct Entry {
  int32_t col;
  int32_t code;
};
}  // namespace

std::string build_symbolic(int n_node, int n_row, int n_quad, const int32_t* cq, int n_beam,
                           const int32_t* cb, int n_known, const int32_t* known, Symbolic& S) {
  if (n_node <= 0 || n_row < 0 || n_row > n_node) return "bad node counts";
  if ((int64_t)n_quad + n_beam >= (1 << 27)) return "too many elements for 27-bit item codes";
  S = Symbolic();
  S.n_node = n_node; S.n_row = n_row; S.n_quad = n_quad; S.n_beam = n_beam;
  for (int64_t i = 0; i < 4LL * n_quad; ++i)
    if (cq[i] < 0 || cq[i] >= n_node) return "quad connectivity out of range";
  for (int64_t i = 0; i < 2LL * n_beam; ++i)
    if (cb[i] < 0 || cb[i] >= n_node) return "beam connectivity out of range";

  // ---- boundary mask (duplicates in `known` are harmless here; the reference's
  // Lagrange system would become singular, model.py:286)
  S.node_mask.assign(n_node, 0);
  for (int i = 0; i < n_known; ++i) {
    if (known[i] < 0 || known[i] >= 6 * n_node) return "known dof out of range";
    S.node_mask[known[i] / 6] |= (uint8_t)(1u << (known[i] % 6));
  }

  // ---- corners per node (gradient gather lists)
  S.node_inc_ptr.assign(n_node + 1, 0);
  for (int e = 0; e < n_beam; ++e)
    for (int a = 0; a < 2; ++a) S.node_inc_ptr[cb[2 * e + a] + 1]++;
  for (int e = 0; e < n_quad; ++e)
    for (int a = 0; a < 4; ++a) S.node_inc_ptr[cq[4 * e + a] + 1]++;
  std::partial_sum(S.node_inc_ptr.begin(), S.node_inc_ptr.end(), S.node_inc_ptr.begin());
  S.node_inc.resize(S.node_inc_ptr[n_node]);
  {
    std::vector<int32_t> pos(S.node_inc_ptr.begin(), S.node_inc_ptr.end() - 1);
    for (int e = 0; e < n_beam; ++e)
      for (int a = 0; a < 2; ++a) S.node_inc[pos[cb[2 * e + a]]++] = ((n_quad + e) << 2) | a;
    for (int e = 0; e < n_quad; ++e)
      for (int a = 0; a < 4; ++a) S.node_inc[pos[cq[4 * e + a]]++] = (e << 2) | a;
  }

  // ---- raw pair items bucketed by block row, in the reference's raw order
  // (diagonal placeholder first so that every row owns its diagonal block)
  std::vector<int64_t> rptr(n_row + 1, 0);
  for (int r = 0; r < n_row; ++r) rptr[r + 1] = 1;
  for (int e = 0; e < n_beam; ++e)
    for (int a = 0; a < 2; ++a)
      if (cb[2 * e + a] < n_row) rptr[cb[2 * e + a] + 1] += 2;
  for (int e = 0; e < n_quad; ++e)
    for (int a = 0; a < 4; ++a)
      if (cq[4 * e + a] < n_row) rptr[cq[4 * e + a] + 1] += 4;
  std::partial_sum(rptr.begin(), rptr.end(), rptr.begin());
  std::vector<Entry> ent(rptr[n_row]);
  {
    std::vector<int64_t> pos(rptr.begin(), rptr.end() - 1);
    for (int r = 0; r < n_row; ++r) ent[pos[r]++] = Entry{r, -1};
    for (int e = 0; e < n_beam; ++e)
      for (int a = 0; a < 2; ++a) {
        const int r = cb[2 * e + a];
        if (r >= n_row) continue;
        for (int b = 0; b < 2; ++b)
          ent[pos[r]++] = Entry{cb[2 * e + b], ((n_quad + e) << 4) | (a << 2) | b};
      }
    for (int e = 0; e < n_quad; ++e)
      for (int a = 0; a < 4; ++a) {
        const int r = cq[4 * e + a];
        if (r >= n_row) continue;
        for (int b = 0; b < 4; ++b) ent[pos[r]++] = Entry{cq[4 * e + b], (e << 4) | (a << 2) | b};
      }
  }

  // ---- per row: stable sort by column -> blocks; contributors keep raw order
  S.rowptr.assign(n_row + 1, 0);
  S.diag_slot.assign(n_row, -1);
  S.colidx.reserve(ent.size() / 2);
  S.blk_row.reserve(ent.size() / 2);
  S.blk_item_ptr.reserve(ent.size() / 2 + 1);
  S.item_code.reserve(ent.size());
  S.blk_item_ptr.push_back(0);
  for (int r = 0; r < n_row; ++r) {
    Entry* b = ent.data() + rptr[r];
    Entry* e = ent.data() + rptr[r + 1];
    std::stable_sort(b, e, [](const Entry& x, const Entry& y) { return x.col < y.col; });
    for (Entry* p = b; p < e;) {
      const int32_t c = p->col;
      if (c == r) S.diag_slot[r] = (int32_t)S.colidx.size();
      S.colidx.push_back(c);
      S.blk_row.push_back(r);
      for (; p < e && p->col == c; ++p)
        if (p->code >= 0) S.item_code.push_back(p->code);
      S.blk_item_ptr.push_back((int32_t)S.item_code.size());
    }
    S.rowptr[r + 1] = (int32_t)S.colidx.size();
  }
  if (S.item_code.size() >= (size_t)INT32_MAX) return "too many pair items";

  // ---- chunks: consecutive blocks with <= kChunkItems items and <= kChunkQuads quads
  const int64_t nnzb = S.nnzb();
  S.item_lel.assign(S.item_code.size(), 0);
  std::vector<int32_t> seen(n_quad, -1), lidx(n_quad, 0);
  S.chunk_blk.push_back(0);
  S.chunk_el_ptr.push_back(0);
  int items = 0, quads = 0, chunk = 0;
  std::vector<int32_t> fresh;
  for (int64_t blk = 0; blk < nnzb; ++blk) {
    const int i0 = S.blk_item_ptr[blk], i1 = S.blk_item_ptr[blk + 1];
    if (i1 - i0 > kChunkItems) return "a block has more contributors than a chunk can hold";
    fresh.clear();
    for (int i = i0; i < i1; ++i) {
      const int el = S.item_code[i] >> 4;
      if (el < n_quad && seen[el] != chunk &&
          std::find(fresh.begin(), fresh.end(), el) == fresh.end())
        fresh.push_back(el);
    }
    if ((int)fresh.size() > kChunkQuads) return "a block touches more quads than a chunk can stage";
    if (items + (i1 - i0) > kChunkItems || quads + (int)fresh.size() > kChunkQuads ||
        blk - S.chunk_blk.back() >= kChunkBlocks) {
      S.chunk_blk.push_back((int32_t)blk);
      S.chunk_el_ptr.push_back((int32_t)S.chunk_els.size());
      ++chunk; items = 0; quads = 0;
    }
    for (int i = i0; i < i1; ++i) {
      const int el = S.item_code[i] >> 4;
      if (el >= n_quad) continue;
      if (seen[el] != chunk) {
        seen[el] = chunk;
        lidx[el] = quads++;
        S.chunk_els.push_back(el);
      }
      S.item_lel[i] = (uint8_t)lidx[el];
    }
    items += i1 - i0;
  }
  S.chunk_blk.push_back((int32_t)nnzb);
  S.chunk_el_ptr.push_back((int32_t)S.chunk_els.size());
  // per chunk: blocks by decreasing contributor count (stable => ascending id within a count)
  S.blk_perm.resize(nnzb);
  std::iota(S.blk_perm.begin(), S.blk_perm.end(), 0);
  for (int c = 0; c + 1 < (int)S.chunk_blk.size(); ++c)
    std::stable_sort(S.blk_perm.begin() + S.chunk_blk[c], S.blk_perm.begin() + S.chunk_blk[c + 1],
                     [&](int32_t x, int32_t y) {
                       return S.blk_item_ptr[x + 1] - S.blk_item_ptr[x] > S.blk_item_ptr[y + 1] - S.blk_item_ptr[y];
                     });
  // ---- per-block boundary word
  S.blk_bc.resize(nnzb);
  for (int64_t blk = 0; blk < nnzb; ++blk) {
    const int r = S.blk_row[blk], c = S.colidx[blk];
    S.blk_bc[blk] = (uint16_t)(S.node_mask[r] | (S.node_mask[c] << 6) | ((r == c) ? (1u << 12) : 0u));
  }

  // ---- warp tasks: runs of consecutive blocks with <= 32 items and <= kTaskQuads distinct quads
  S.item_desc.assign(S.item_code.size(), 0);
  S.tasks_ok = true;
  {
    std::fill(seen.begin(), seen.end(), -1);
    int task = 0, t_items = 0, t_quads = 0, t_blks = 0;
    int32_t t_blk0 = 0, t_item0 = 0, t_el0 = 0;
    auto close_task = [&](int64_t blk_end) {
      if (blk_end > t_blk0) {
        S.task_meta.push_back(t_blk0); S.task_meta.push_back(t_item0); S.task_meta.push_back(t_el0);
        S.task_meta.push_back(t_blks | (t_items << 8) | (t_quads << 16));
        ++task;
      }
      t_blk0 = (int32_t)blk_end; t_item0 = S.blk_item_ptr[blk_end]; t_el0 = (int32_t)S.task_els.size();
      t_items = t_quads = t_blks = 0;
    };
    for (int64_t blk = 0; blk < nnzb && S.tasks_ok; ++blk) {
      const int i0 = S.blk_item_ptr[blk], i1 = S.blk_item_ptr[blk + 1];
      fresh.clear();
      for (int i = i0; i < i1; ++i) {
        const int el = S.item_code[i] >> 4;
        if (el < n_quad && seen[el] != task && std::find(fresh.begin(), fresh.end(), el) == fresh.end())
          fresh.push_back(el);
      }
      if (t_items + (i1 - i0) > kTaskItems || t_quads + (int)fresh.size() > kTaskQuads || t_blks >= 32) {
        close_task(blk);
        // the quads seen by the closed task are fresh again for the new one
        fresh.clear();
        for (int i = i0; i < i1; ++i) {
          const int el = S.item_code[i] >> 4;
          if (el < n_quad && std::find(fresh.begin(), fresh.end(), el) == fresh.end()) fresh.push_back(el);
        }
      }
      // a block one warp cannot hold, or an empty block (a node without elements): chunked path
      if (i1 == i0 || i1 - i0 > kTaskItems || (int)fresh.size() > kTaskQuads) { S.tasks_ok = false; break; }
      for (int i = i0; i < i1; ++i) {
        const int code = S.item_code[i];
        const int el = code >> 4;
        unsigned d = (unsigned)t_blks | ((unsigned)(code & 15) << 5);   // (a<<2|b) -> bits 5..8: b in 5-6, a in 7-8
        if (el >= n_quad) {
          d |= kDescBeam;
        } else {
          if (seen[el] != task) {
            seen[el] = task;
            lidx[el] = t_quads++;
            S.task_els.push_back(el);
          }
          d |= (unsigned)lidx[el] << 9;
        }
        if (i == i0) d |= kDescFirst;
        S.item_desc[i] = (uint16_t)d;
      }
      t_items += i1 - i0;
      ++t_blks;
    }
    if (S.tasks_ok) close_task(nnzb);
    if (!S.tasks_ok) { S.task_meta.clear(); S.task_els.clear(); }
  }
  return "";
}

std::string build_symbolic_from_bsr(int n_node, const int32_t* rowptr, const int32_t* colidx, int n_known,
                                    const int32_t* known, Symbolic& S) {
  if (n_node <= 0 || !rowptr || rowptr[0] != 0) return "bad block-CSR pattern";
  S = Symbolic();
  S.n_node = n_node; S.n_row = n_node;
  S.node_mask.assign(n_node, 0);
  for (int i = 0; i < n_known; ++i) {
    if (known[i] < 0 || known[i] >= 6 * n_node) return "known dof out of range";
    S.node_mask[known[i] / 6] |= (uint8_t)(1u << (known[i] % 6));
  }
  S.rowptr.assign(rowptr, rowptr + n_node + 1);
  const int64_t nnzb = rowptr[n_node];
  S.colidx.assign(colidx, colidx + nnzb);
  S.blk_row.resize(nnzb);
  S.diag_slot.assign(n_node, -1);
  for (int r = 0; r < n_node; ++r) {
    if (rowptr[r + 1] < rowptr[r]) return "row pointers not monotone";
    for (int s = rowptr[r]; s < rowptr[r + 1]; ++s) {
      const int c = colidx[s];
      if (c < 0 || c >= n_node) return "column index out of range";
      if (s > rowptr[r] && colidx[s - 1] >= c) return "columns must be sorted and unique within a row";
      S.blk_row[s] = r;
      if (c == r) S.diag_slot[r] = s;
    }
    if (S.diag_slot[r] < 0) return "every block row needs its diagonal block";
  }
  S.blk_item_ptr.assign(nnzb + 1, 0);
  S.blk_perm.resize(nnzb);
  for (int64_t i = 0; i < nnzb; ++i) S.blk_perm[i] = (int32_t)i;
  S.chunk_blk = {0, (int32_t)nnzb};
  S.chunk_el_ptr = {0, 0};
  S.node_inc_ptr.assign(n_node + 1, 0);
  return "";
}

int mg_aggregate(int n, const int32_t* rowptr, const int32_t* colidx, int32_t* agg) {
  std::fill(agg, agg + n, -1);
  int na = 0;
  for (int i = 0; i < n; ++i) {
    if (agg[i] >= 0) continue;
    bool all_free = true;
    for (int s = rowptr[i]; s < rowptr[i + 1]; ++s)
      if (agg[colidx[s]] >= 0) { all_free = false; break; }
    if (!all_free) continue;
    for (int s = rowptr[i]; s < rowptr[i + 1]; ++s) agg[colidx[s]] = na;
    agg[i] = na;
    ++na;
  }
  for (int i = 0; i < n; ++i) {
    if (agg[i] >= 0) continue;
    int got = -1;
    for (int s = rowptr[i]; s < rowptr[i + 1]; ++s)
      if (agg[colidx[s]] >= 0) { got = agg[colidx[s]]; break; }
    agg[i] = (got >= 0) ? got : na++;
  }
  return na;
}

int64_t mg_pattern_lists(int64_t m, const int32_t* row, const int32_t* col, const int32_t* left,
                         const int32_t* right, int n_row, int32_t* rowptr, int32_t* ocol, int32_t* ptr,
                         int32_t* left_o, int32_t* right_o) {
  std::vector<int64_t> start(n_row + 1, 0);
  for (int64_t t = 0; t < m; ++t) start[row[t] + 1]++;
  for (int r = 0; r < n_row; ++r) start[r + 1] += start[r];
  std::vector<int64_t> order(m);
  {
    std::vector<int64_t> pos(start.begin(), start.end() - 1);
    for (int64_t t = 0; t < m; ++t) order[pos[row[t]]++] = t;      // stable by row
  }
  int64_t nb = 0;
  rowptr[0] = 0;
  for (int r = 0; r < n_row; ++r) {
    int64_t* b = order.data() + start[r];
    int64_t* e = order.data() + start[r + 1];
    std::stable_sort(b, e, [&](int64_t x, int64_t y) { return col[x] < col[y]; });
    for (int64_t* p = b; p < e; ++p) {
      const int64_t o = p - order.data();
      if (p == b || col[*p] != col[*(p - 1)]) {
        ocol[nb] = col[*p];
        ptr[nb] = (int32_t)o;
        ++nb;
      }
      left_o[o] = left[*p];
      right_o[o] = right[*p];
    }
    rowptr[r + 1] = (int32_t)nb;
  }
  ptr[nb] = (int32_t)m;
  return nb;
}

}  // namespace jsso
