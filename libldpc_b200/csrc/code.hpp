// Host-side code model and device-layout builder.
//
// HostCode   : what the reference's ldpc_code + sparse_csr hold (src/core/ldpc.cpp:40-101,
//              src/core/sparse.h:91-153): edge list in FILE ORDER, per-row / per-column edge lists in
//              file order (they fix the box-plus recursion order and the variable-node summation
//              order), puncture/shorten lists, bit_pos, max degree.
// TileLayout : frame-in-lane mapping used by the erasure decoder (bec_kernel.cuh): frames in the fast
//              dimension, degree-sorted node groups balanced over warps, edge-slot-major message slots.
// SegLayout  : the vector-tile / segment mapping of the BP and min-sum decode kernel (tile4.cuh).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace b200
{
    struct HostCode
    {
        int nc = 0, mc = 0, nnz = 0;
        std::vector<int> e_row, e_col;         // [nnz] file order
        std::vector<int> row_ptr, row_edge;    // rowN (edge ids, file order within the row)
        std::vector<int> col_ptr, col_edge;    // colN (edge ids, file order within the column)
        std::vector<int> puncture, shorten;    // header lists, as given
        std::vector<int> bit_pos;              // transmitted positions, ascending
        int max_degree = 0, max_cn_degree = 0, max_vn_degree = 0, min_cn_degree = 0;
        int kc() const { return nc - mc; }
        int nct() const { return (int)bit_pos.size(); }
        int mct() const { return mc - (int)puncture.size(); }
        int kct() const { return nct() - mct(); }
        bool empty() const { return nnz == 0; }

        // Parses a code file.  parse_header=true applies the leading "key: value" header rule
        // (ldpc.cpp:49-76); false reads every line as an edge (generator matrix, ldpc.cpp:103-106).
        // Throws std::runtime_error("can not open file for reading") like sparse.h:99.
        void load(const std::string &path, bool parse_header);

        // GF(2) helpers (sparse.h:162-218, 227-294)
        void multiply_left(const uint8_t *left, uint8_t *result_accum) const;  // result[col] ^= left[row]
        void multiply_right(const uint8_t *right, uint8_t *result_accum) const; // result[row] ^= right[col]
        int rank() const;
    };

    // One scheduled node entry: first slot, slot stride, degree.
    constexpr uint32_t IDLE_NODE = 0xFFFFFFFFu;

    struct TileLayout
    {
        int fpc = 0;       // frames per CTA (power of two <= 32)
        int threads = 0;   // threads per CTA
        int nt = 0;        // node threads = threads / fpc
        int npw = 0;       // node threads per warp = 32 / fpc
        int n_slots = 0;   // padded number of message slots
        int cn_rounds = 0, vn_rounds = 0;
        // check side: entry k = round * nt + node_thread
        std::vector<uint32_t> cn_desc;  // first slot (bits 0-23) | degree (bits 24-31); IDLE_NODE = none
        std::vector<uint32_t> cn_col;   // [n_slots] variable id gathered by each slot (0 for padding)
        // variable side
        std::vector<uint32_t> vn_desc;  // first index into vn_slot (bits 0-22) | degree (23-30) | transmitted (31); IDLE_NODE = none
        std::vector<uint32_t> vn_id;    // variable id (0 for idle entries)
        std::vector<uint32_t> vn_slot;  // [n_vslots] message slot of the k-th edge (file order) at q0 + k*npw
        int n_vslots = 0;
        std::vector<int> edge_slot;     // [nnz] file-order edge -> message slot (for tests / debugging)

        void build(const HostCode &code, int fpc, int threads);
    };

    // Segment mapping (tile4.cuh).  One thread owns ONE 16-byte vector of adjacent frame lanes (2 doubles /
    // 4 floats) of one node; `lanes` warp lanes serve the same node (a CTA holds lanes * 16/sizeof(T)
    // frames); npw = 32/lanes nodes of equal degree form a warp task.  Tasks are spread longest-first over
    // the warps and each warp's task list is stored run-length encoded as SEGMENTS (runs of tasks of one
    // degree); message slots, variable positions ("positions": variables renumbered into schedule order so
    // that the variable phase walks llr/out contiguously) and index entries are numbered warp-major in
    // list order.  Inside a segment a warp therefore only advances three pointers by compile-time
    // strides: no per-task descriptor, no per-task dispatch.
    //
    //   message slot  = seg.slot_base + (t*deg + k)*npw + j      (task t of the segment, edge k, node j)
    //   position      = seg.pos_base + t*npw + j                 (variable side)
    //   index entries = seg.idx_base + (t*npw + j)*stride + k*isz bytes while a node's deg entries fit 8
    //                   bytes (`stride` = idx_stride(deg, isz)); longer blocks are cut into 16-byte chunks
    //                   stored chunk-major, seg.idx_base + t*npw*stride + (k/epc)*npw*16 + j*16 + (k%epc)*isz
    //                   with epc = 16/isz, so a thread fetches 16 bytes' worth of its node's indices with one
    //                   vector load and a warp's load of a chunk is contiguous (bank-conflict free).
    //                   Nodes missing from a ragged last task keep all-zero entries and own padded slots /
    //                   positions, so their threads can run the node update unconditionally.
    //                   Check side: entry = position gathered by edge k; variable side: entry = message
    //                   slot of edge k (both in file order).  Entries are stored PRE-SCALED to the byte
    //                   offset of the record they name (index * 16 * lanes; in 16-byte units when isz = 2).
    struct SegLayout
    {
        int lanes = 0, threads = 0, warps = 0, npw = 0, isz = 0; // isz = bytes per index entry (2 or 4)
        int n_slots = 0, n_pos = 0;
        int cn_max_segs = 0, vn_max_segs = 0;
        // segment s of warp w at 4*(w*max_segs + s): {degree | nodes per task << 8 | tasks << 16,
        //   byte offset of the first message slot (check side) / first position (variable side) record,
        //   idx byte offset, 0}; first word 0 terminates
        std::vector<uint32_t> cn_seg, vn_seg;
        std::vector<uint8_t> cn_idx, vn_idx; // packed index entries (isz bytes each), 16-byte aligned per segment
        std::vector<uint32_t> var_pos;       // [nc] variable id -> position
        std::vector<int> edge_slot;          // [nnz] file-order edge -> message slot
        long cn_path = 0, vn_path = 0, cn_work = 0, vn_work = 0; // longest warp list / total, in edge steps (diagnostics)

        static int idx_stride(int deg, int isz)
        {
            const int b = deg * isz;
            return b <= 2 ? 2 : b <= 4 ? 4 : b <= 8 ? 8 : (b + 15) & ~15;
        }
        void build(const HostCode &code, int lanes, int threads, int isz);
    };

    // Message slots of the bit-sliced erasure kernel (bec_slice.cuh): a proper 32-colouring of the code's edges seen as
    // the edges of the bipartite multigraph of warp access groups, slot % 32 = colour (code.cpp).
    struct BecSliceLayout
    {
        int n_slots = 0;                       // 32 * size of the largest colour class
        std::vector<uint16_t> row_slot, col_slot; // slots in the order of HostCode::row_edge / col_edge
        std::vector<int> edge_slot;            // [nnz] file-order edge -> slot
        void build(const HostCode &code);
    };

    // Layered schedule (layered.cuh).  `layers` partitions the checks; no two checks of a layer may share a variable.
    // Per (layer, warp) a run-length list of warp tasks (npw checks of equal degree each); message slots are numbered in list
    // order, slot of edge k of node j of task t = segment base + (t*deg + k)*npw + j; idx[slot] = byte offset of the posterior
    // record the edge gathers (variable id * 16 * lanes).
    struct LayeredLayout
    {
        int lanes = 0, threads = 0, warps = 0, npw = 0, n_layers = 0, max_segs = 0, n_slots = 0;
        std::vector<uint32_t> seg; // [(layer*warps + warp)*max_segs + s][4] = {degree | nodes << 8 | tasks << 16, first slot, 0, 0}
        std::vector<uint32_t> idx; // [n_slots]
        std::vector<int> edge_slot;
        void build(const HostCode &code, const std::vector<std::vector<int>> &layers, int lanes, int threads);
    };
    // first-fit layers: check i joins the first layer none of whose checks shares a variable with it
    std::vector<std::vector<int>> auto_layers(const HostCode &code);
    // legacy layer file (gpu/ldpc/ldpc.cpp:111-138): "nl: N", then per layer "cn[i]: W" followed by W check indices
    std::vector<std::vector<int>> read_layer_file(const std::string &path);
    // throws unless `layers` is a partition of the checks in which no layer holds two checks sharing a variable
    void validate_layers(const HostCode &code, const std::vector<std::vector<int>> &layers);
} // namespace b200
