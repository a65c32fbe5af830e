// Symbolic pass (host, once per model): everything that is a pure function of
// the connectivity and the boundary conditions.
//
// The reference recomputes COO indices per call (JaxSSO/element.py:141-149,
// 1097-1106) and lexicographically sorts + segment-sums all 144 n_b + 576 n_q raw
// entries on every evaluation (JaxSSO/assemblemodel.py:156-162).  Its
// sorted-unique (row, col) set is exactly the 6x6 block expansion of the node
// adjacency graph, which is what this pass builds once as block-CSR, together
// with the contributor lists that let the numeric assembly sum duplicates in a
// fixed order without atomics.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace jsso {

constexpr int kChunkBlocks = 128; // block slots per assembly chunk (= CTA size: one thread per block)
constexpr int kChunkItems = 224;  // pair items per assembly chunk
constexpr int kChunkQuads = 36;   // distinct quads whose geometry a chunk stages in shared memory
// warp tasks of the two-kernel assembly (quad_geometry_kernel + assemble_tasks_kernel): one warp sums
// <= 32 pair items (one per lane) of a run of consecutive block slots
constexpr int kTaskItems = 32;    // = warp size
constexpr int kTaskQuads = 8;     // distinct quads whose records one warp stages
// item descriptor bits (uint16)
constexpr unsigned kDescBeam = 1u << 12, kDescFirst = 1u << 13;   // lb:0-4  b:5-6  a:7-8  lq:9-11

struct Symbolic {
  int n_node = 0, n_row = 0, n_quad = 0, n_beam = 0;
  // block-CSR over nodes; rows = owned nodes [0, n_row), columns = all local nodes
  std::vector<int32_t> rowptr, colidx, blk_row, diag_slot;
  // pair items (element, a, b) sorted by block slot, beams before quads, then by
  // element id (the reference's concatenation order, assemblemodel.py:202-211)
  std::vector<int32_t> blk_item_ptr;  // nnzb + 1
  std::vector<int32_t> item_code;     // (elem << 4) | (a << 2) | b; beams use elem = n_quad + id
  std::vector<uint8_t> item_lel;      // quad items: index into the chunk's staged-geometry list
  std::vector<int32_t> chunk_blk;     // n_chunk + 1 block boundaries
  std::vector<int32_t> blk_perm;      // nnzb: within each chunk, its block ids ordered by decreasing
                                      // contributor count (warps then loop a uniform number of times)
  std::vector<int32_t> chunk_el_ptr;  // n_chunk + 1
  std::vector<int32_t> chunk_els;     // quads staged per chunk
  // warp tasks (see kTaskItems): task_meta = 4 ints per task {blk0, item0, el0, n_blk | n_item<<8 | n_el<<16};
  // task_els = quads staged per task; item_desc per pair item (same order as item_code): local block,
  // node pair, local quad, beam flag, first-item-of-block flag; blk_bc per block = row mask | col mask<<6 |
  // diagonal<<12.  tasks_ok is false when some block cannot be handled by one warp (> 32 contributors or
  // > kTaskQuads quads): the chunked single-kernel path is used instead.
  std::vector<int32_t> task_meta, task_els;
  std::vector<uint16_t> item_desc, blk_bc;
  bool tasks_ok = false;
  int n_task() const { return (int)(task_meta.size() / 4); }
  // per node: incident (element, local node) corners, for the gradient gather
  std::vector<int32_t> node_inc_ptr;  // n_node + 1
  std::vector<int32_t> node_inc;      // (elem << 2) | a; beams use elem = n_quad + id
  std::vector<uint8_t> node_mask;     // bit k set <=> dof 6*node+k is prescribed (zero)
  int64_t nnzb() const { return (int64_t)colidx.size(); }
  int64_t n_items() const { return (int64_t)item_code.size(); }
  int n_chunk() const { return (int)chunk_blk.size() - 1; }
};

// Returns an empty string on success, else an error message.
std::string build_symbolic(int n_node, int n_row, int n_quad, const int32_t* cnct_quads, int n_beam,
                           const int32_t* cnct_beams, int n_known, const int32_t* known,
                           Symbolic& out);

// Pattern-only symbolic state (solver-plugin compatibility mode): the block-CSR pattern is given,
// there are no element contributor lists.
std::string build_symbolic_from_bsr(int n_node, const int32_t* rowptr, const int32_t* colidx, int n_known,
                                    const int32_t* known, Symbolic& out);

// Greedy aggregation of the smoothed-aggregation multigrid (Vanek et al.), the host loop of
// jaxsso_b200/multigrid.py::aggregate_py restated in C++ (identical, deterministic result): a node whose whole
// neighbourhood is free roots an aggregate of itself + neighbours; leftovers join the aggregate of their first
// aggregated neighbour (column order) or become singletons.  Returns the number of aggregates.
int mg_aggregate(int n, const int32_t* rowptr, const int32_t* colidx, int32_t* agg);

// Gather-list construction of the multigrid symbolic setup: m triples (row, col, left slot, right slot) ->
// block pattern sorted by (row, col) (rowptr[n_row+1], ocol[n_blk]) and, per output block, its (left, right)
// pairs IN INPUT ORDER (ptr[n_blk+1], left_o[m], right_o[m]).  Stable two-level sort (counting sort by row,
// stable sort by column inside a row).  Output arrays are sized for the worst case n_blk = m.  Returns n_blk.
int64_t mg_pattern_lists(int64_t m, const int32_t* row, const int32_t* col, const int32_t* left,
                         const int32_t* right, int n_row, int32_t* rowptr, int32_t* ocol, int32_t* ptr,
                         int32_t* left_o, int32_t* right_o);

}  // namespace jsso
