#include "code.hpp"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <numeric>
#include <queue>
#include <stdexcept>

namespace b200
{
    namespace
    {
        // whitespace separated integers until the first token that is not one ("record >> index" loop)
        std::vector<int> leading_ints(const char *s)
        {
            std::vector<int> v;
            for (;;)
            {
                char *end = nullptr;
                long x = std::strtol(s, &end, 10);
                if (end == s) break;
                v.push_back(static_cast<int>(x));
                s = end;
            }
            return v;
        }
    } // namespace

    // Semantics follow src/core/ldpc.cpp:40-101 and src/core/sparse.h:91-153:
    //  * leading lines containing ':' form the header; only keys containing "puncture"/"shorten" matter
    //  * every following line is "row col [value]"; any listed entry is a one; dimensions = max index + 1
    //  * edge ids are file positions; per-node adjacency keeps file order
    // Lines that do not start with two integers (blank/trailing lines) are ignored; the reference pushes
    // uninitialised indices for those (undefined behaviour), so there is nothing to be compatible with.
    void HostCode::load(const std::string &path, bool parse_header)
    {
        std::ifstream in(path);
        if (!in.good()) throw std::runtime_error("can not open file for reading");
        *this = HostCode();
        std::string line;
        bool header = parse_header;
        int max_r = 0, max_c = 0;
        while (std::getline(in, line))
        {
            if (header)
            {
                const auto colon = line.find(':');
                if (colon != std::string::npos)
                {
                    const std::string key = line.substr(0, colon);
                    if (key.find("puncture") != std::string::npos)
                    {
                        auto v = leading_ints(line.c_str() + colon + 1);
                        puncture.insert(puncture.end(), v.begin(), v.end());
                    }
                    else if (key.find("shorten") != std::string::npos)
                    {
                        auto v = leading_ints(line.c_str() + colon + 1);
                        shorten.insert(shorten.end(), v.begin(), v.end());
                    }
                    continue;
                }
                header = false;
            }
            auto v = leading_ints(line.c_str());
            if (v.size() < 2) continue;
            if (v[0] < 0 || v[1] < 0) throw std::runtime_error("negative index in code file");
            e_row.push_back(v[0]);
            e_col.push_back(v[1]);
            max_r = std::max(max_r, v[0]);
            max_c = std::max(max_c, v[1]);
        }
        nnz = static_cast<int>(e_row.size());
        mc = max_r + 1;
        nc = max_c + 1;

        row_ptr.assign(mc + 1, 0);
        col_ptr.assign(nc + 1, 0);
        for (int e = 0; e < nnz; ++e) { ++row_ptr[e_row[e] + 1]; ++col_ptr[e_col[e] + 1]; }
        std::partial_sum(row_ptr.begin(), row_ptr.end(), row_ptr.begin());
        std::partial_sum(col_ptr.begin(), col_ptr.end(), col_ptr.begin());
        row_edge.assign(nnz, 0);
        col_edge.assign(nnz, 0);
        {
            std::vector<int> rf(row_ptr.begin(), row_ptr.end() - 1), cf(col_ptr.begin(), col_ptr.end() - 1);
            for (int e = 0; e < nnz; ++e) { row_edge[rf[e_row[e]]++] = e; col_edge[cf[e_col[e]]++] = e; }
        }
        max_cn_degree = max_vn_degree = 0;
        min_cn_degree = nnz;
        for (int i = 0; i < mc; ++i)
        {
            max_cn_degree = std::max(max_cn_degree, row_ptr[i + 1] - row_ptr[i]);
            min_cn_degree = std::min(min_cn_degree, row_ptr[i + 1] - row_ptr[i]);
        }
        for (int i = 0; i < nc; ++i) max_vn_degree = std::max(max_vn_degree, col_ptr[i + 1] - col_ptr[i]);
        max_degree = std::max(max_cn_degree, max_vn_degree);

        std::vector<char> removed(nc, 0);
        for (int p : puncture) if (p >= 0 && p < nc) removed[p] = 1;
        for (int s : shorten) if (s >= 0 && s < nc) removed[s] = 1;
        for (int i = 0; i < nc; ++i) if (!removed[i]) bit_pos.push_back(i);
    }

    void HostCode::multiply_left(const uint8_t *left, uint8_t *result) const
    {
        for (int e = 0; e < nnz; ++e) result[e_col[e]] ^= (left[e_row[e]] & 1);
    }

    void HostCode::multiply_right(const uint8_t *right, uint8_t *result) const
    {
        for (int e = 0; e < nnz; ++e) result[e_row[e]] ^= (right[e_col[e]] & 1);
    }

    // GF(2) rank by bit-packed Gauss elimination over 64-bit words (same value as the list-based
    // elimination of sparse.h:227-294).
    int HostCode::rank() const
    {
        const size_t words = (static_cast<size_t>(nc) + 63) / 64;
        std::vector<uint64_t> m(static_cast<size_t>(mc) * words, 0);
        for (int e = 0; e < nnz; ++e) m[e_row[e] * words + (e_col[e] >> 6)] ^= 1ull << (e_col[e] & 63);
        int r = 0;
        for (int c = 0; c < nc && r < mc; ++c)
        {
            const size_t w = c >> 6;
            const uint64_t bit = 1ull << (c & 63);
            int piv = -1;
            for (int i = r; i < mc; ++i) if (m[i * words + w] & bit) { piv = i; break; }
            if (piv < 0) continue;
            if (piv != r) std::swap_ranges(m.begin() + piv * words, m.begin() + (piv + 1) * words, m.begin() + r * words);
            const uint64_t *src = &m[r * words];
            for (int i = r + 1; i < mc; ++i)
            {
                uint64_t *dst = &m[i * words];
                if (dst[w] & bit) for (size_t k = w; k < words; ++k) dst[k] ^= src[k];
            }
            ++r;
        }
        return r;
    }

    namespace
    {
        struct Group
        {
            int degree;
            std::vector<int> nodes; // <= npw node ids
        };

        // Same-degree nodes are packed npw at a time (one warp executes one group per round with all its
        // node threads on equal trip counts); groups are then spread over warps longest-first.
        // side: 0 = historical cost model (degree + 3); 1 = check tasks, 2 = variable tasks of the segment kernel, whose
        // costs follow the measured instruction counts of the node bodies (a check update is 3*deg - 4 pairwise
        // operations plus loads/stores, a variable update one gather + add per edge plus a fixed part).
        std::vector<std::vector<Group>> schedule(const std::vector<int> &degree, int npw, int warps, int side = 0)
        {
            auto cost = [side](int deg) -> long
            {
                if (side == 1) return 10L * std::max(3 * deg - 4, 1) + 20;
                if (side == 2) return 4L * deg + 12;
                return deg + 3;
            };
            std::vector<int> order(degree.size());
            std::iota(order.begin(), order.end(), 0);
            std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return degree[a] > degree[b]; });
            std::vector<Group> groups;
            for (size_t i = 0; i < order.size();)
            {
                Group g;
                g.degree = degree[order[i]];
                while (i < order.size() && degree[order[i]] == g.degree && (int)g.nodes.size() < npw) g.nodes.push_back(order[i++]);
                groups.push_back(std::move(g));
            }
            using Load = std::pair<long, int>; // (load, warp)
            std::priority_queue<Load, std::vector<Load>, std::greater<Load>> heap;
            for (int w = 0; w < warps; ++w) heap.push({0, w});
            std::vector<std::vector<Group>> per_warp(warps);
            for (auto &g : groups)
            {
                auto [load, w] = heap.top();
                heap.pop();
                per_warp[w].push_back(g);
                heap.push({load + cost(g.degree), w});
            }
            return per_warp;
        }
    } // namespace

    void TileLayout::build(const HostCode &code, int fpc_, int threads_)
    {
        if (fpc_ < 1 || fpc_ > 32 || (fpc_ & (fpc_ - 1))) throw std::runtime_error("frames_per_cta must be a power of two <= 32");
        if (threads_ < 32 || threads_ > 1024 || threads_ % 32) throw std::runtime_error("threads_per_cta must be a multiple of 32 <= 1024");
        if (code.nnz >= (1 << 22)) throw std::runtime_error("code too large (nnz >= 2^22)");
        if (code.max_cn_degree > 64) throw std::runtime_error("check degree > 64 not supported");
        if (code.max_vn_degree > 255) throw std::runtime_error("variable degree > 255 not supported");
        if (code.min_cn_degree < 2) throw std::runtime_error("check nodes of degree < 2 are not supported (undefined in the reference)");
        fpc = fpc_;
        threads = threads_;
        npw = 32 / fpc;
        nt = threads / fpc;
        const int warps = threads / 32;

        std::vector<int> cdeg(code.mc), vdeg(code.nc);
        for (int i = 0; i < code.mc; ++i) cdeg[i] = code.row_ptr[i + 1] - code.row_ptr[i];
        for (int i = 0; i < code.nc; ++i) vdeg[i] = code.col_ptr[i + 1] - code.col_ptr[i];
        std::vector<char> tx(code.nc, 0);
        for (int p : code.bit_pos) tx[p] = 1;

        // ---- check side -------------------------------------------------------------------
        auto cs = schedule(cdeg, npw, warps);
        cn_rounds = 0;
        for (auto &l : cs) cn_rounds = std::max<int>(cn_rounds, (int)l.size());
        cn_desc.assign((size_t)cn_rounds * nt, IDLE_NODE);
        edge_slot.assign(code.nnz, -1);
        cn_col.clear();
        int base = 0;
        for (int r = 0; r < cn_rounds; ++r)
            for (int w = 0; w < warps; ++w)
            {
                if (r >= (int)cs[w].size()) continue;
                const Group &g = cs[w][r];
                cn_col.resize(base + (size_t)npw * g.degree, 0);
                for (int j = 0; j < (int)g.nodes.size(); ++j)
                {
                    const int row = g.nodes[j];
                    cn_desc[(size_t)r * nt + w * npw + j] = (uint32_t)(base + j) | ((uint32_t)g.degree << 24);
                    for (int k = 0; k < g.degree; ++k)
                    {
                        const int e = code.row_edge[code.row_ptr[row] + k];
                        const int slot = base + k * npw + j;
                        edge_slot[e] = slot;
                        cn_col[slot] = (uint32_t)code.e_col[e];
                    }
                }
                base += npw * g.degree;
            }
        n_slots = base;

        // ---- variable side ----------------------------------------------------------------
        auto vs = schedule(vdeg, npw, warps);
        vn_rounds = 0;
        for (auto &l : vs) vn_rounds = std::max<int>(vn_rounds, (int)l.size());
        vn_desc.assign((size_t)vn_rounds * nt, IDLE_NODE);
        vn_id.assign((size_t)vn_rounds * nt, 0);
        vn_slot.clear();
        int qbase = 0;
        for (int r = 0; r < vn_rounds; ++r)
            for (int w = 0; w < warps; ++w)
            {
                if (r >= (int)vs[w].size()) continue;
                const Group &g = vs[w][r];
                vn_slot.resize(qbase + (size_t)npw * g.degree, 0);
                for (int j = 0; j < (int)g.nodes.size(); ++j)
                {
                    const int col = g.nodes[j];
                    const size_t k0 = (size_t)r * nt + w * npw + j;
                    vn_id[k0] = (uint32_t)col;
                    vn_desc[k0] = (uint32_t)(qbase + j) | ((uint32_t)g.degree << 23) | (tx[col] ? 0x80000000u : 0u);
                    for (int k = 0; k < g.degree; ++k)
                        vn_slot[qbase + k * npw + j] = (uint32_t)edge_slot[code.col_edge[code.col_ptr[col] + k]];
                }
                qbase += npw * g.degree;
            }
        n_vslots = qbase;
        if (n_slots >= (1 << 23) || n_vslots >= (1 << 23)) throw std::runtime_error("layout too large");
    }

    void SegLayout::build(const HostCode &code, int lanes_, int threads_, int isz_)
    {
        if (lanes_ < 1 || lanes_ > 4 || (lanes_ & (lanes_ - 1))) throw std::runtime_error("lanes per node must be 1, 2 or 4");
        if (threads_ < 32 || threads_ > 1024 || threads_ % 32) throw std::runtime_error("threads_per_cta must be a multiple of 32 <= 1024");
        if (isz_ != 2 && isz_ != 4) throw std::runtime_error("index entries are 2 or 4 bytes");
        if (code.nnz >= (1 << 22)) throw std::runtime_error("code too large (nnz >= 2^22)");
        if (code.max_cn_degree > 64) throw std::runtime_error("check degree > 64 not supported");
        if (code.max_vn_degree > 255) throw std::runtime_error("variable degree > 255 not supported");
        if (code.min_cn_degree < 2) throw std::runtime_error("check nodes of degree < 2 are not supported (undefined in the reference)");
        lanes = lanes_;
        threads = threads_;
        warps = threads / 32;
        npw = 32 / lanes;
        isz = isz_;

        std::vector<int> cdeg(code.mc), vdeg(code.nc);
        for (int i = 0; i < code.mc; ++i) cdeg[i] = code.row_ptr[i + 1] - code.row_ptr[i];
        for (int i = 0; i < code.nc; ++i) vdeg[i] = code.col_ptr[i + 1] - code.col_ptr[i];

        struct Seg { int deg, cnt, ntasks; uint32_t base, idx; size_t first; };
        auto encode = [&](const std::vector<Group> &list, std::vector<Seg> &segs)
        { // run-length encodes one warp's task list; a task with fewer than npw nodes gets its own segment
            for (size_t i = 0; i < list.size();)
            {
                const int deg = list[i].degree, cnt = (int)list[i].nodes.size();
                size_t n = 1;
                while (i + n < list.size() && list[i + n].degree == deg && (int)list[i + n].nodes.size() == cnt && n < 65535) ++n;
                segs.push_back({deg, cnt, (int)n, 0u, 0u, i});
                i += n;
            }
        };
        // entries are stored pre-scaled: byte offset of the record (index * 16 * lanes) as uint32, or that offset
        // in 16-byte units (index * lanes) as uint16
        auto put = [&](std::vector<uint8_t> &buf, size_t off, uint32_t v)
        {
            v *= (uint32_t)(isz == 2 ? lanes : 16 * lanes);
            if (isz == 2)
            {
                if (v > 0xFFFFu) throw std::runtime_error("index does not fit 16 bits");
                buf[off] = (uint8_t)(v & 0xFF); buf[off + 1] = (uint8_t)(v >> 8);
            }
            else
            {
                buf[off] = (uint8_t)(v & 0xFF); buf[off + 1] = (uint8_t)((v >> 8) & 0xFF);
                buf[off + 2] = (uint8_t)((v >> 16) & 0xFF); buf[off + 3] = (uint8_t)(v >> 24);
            }
        };
        // byte offset of entry k of node j of task t inside a segment's index block: node-major while a node's
        // entries fit 8 bytes, else 16-byte chunks stored chunk-major ([task][chunk][node]) so that the
        // vector load of a chunk is contiguous over the nodes of a warp task (bank-conflict free)
        auto idx_offset = [&](int t, int j, int k, int stride) -> size_t
        {
            if (stride < 16) return ((size_t)t * npw + j) * stride + (size_t)k * isz;
            const int epc = 16 / isz;
            return (size_t)t * npw * stride + (size_t)(k / epc) * npw * 16 + (size_t)j * 16 + (size_t)(k % epc) * isz;
        };
        auto flatten = [&](const std::vector<std::vector<Seg>> &segs, int &max_segs, std::vector<uint32_t> &out)
        {
            max_segs = 1;
            for (auto &l : segs) max_segs = std::max<int>(max_segs, (int)l.size() + 1); // + terminator
            out.assign((size_t)4 * warps * max_segs, 0);
            for (int w = 0; w < warps; ++w)
                for (size_t s = 0; s < segs[w].size(); ++s)
                {
                    const Seg &g = segs[w][s];
                    uint32_t *d = &out[4 * ((size_t)w * max_segs + s)];
                    d[0] = (uint32_t)g.deg | ((uint32_t)g.cnt << 8) | ((uint32_t)g.ntasks << 16);
                    d[1] = g.base * (uint32_t)(16 * lanes); // byte offset of the first record
                    d[2] = g.idx;
                }
        };

        // ---- variable side first: it defines the positions the check side gathers from -------------
        auto vs = schedule(vdeg, npw, warps, 2);
        std::vector<std::vector<Seg>> vsegs(warps), csegs(warps);
        var_pos.assign(code.nc, 0);
        size_t pbase = 0, ibase = 0;
        vn_path = vn_work = 0;
        for (int w = 0; w < warps; ++w)
        {
            encode(vs[w], vsegs[w]);
            long path = 0;
            for (Seg &sg : vsegs[w])
            {
                ibase = (ibase + 15) & ~(size_t)15;
                sg.base = (uint32_t)pbase;
                sg.idx = (uint32_t)ibase;
                for (int t = 0; t < sg.ntasks; ++t)
                {
                    const Group &g = vs[w][sg.first + t];
                    for (int j = 0; j < (int)g.nodes.size(); ++j) var_pos[g.nodes[j]] = (uint32_t)(pbase + j);
                    pbase += npw;
                }
                ibase += (size_t)sg.ntasks * npw * idx_stride(sg.deg, isz);
                path += (long)sg.ntasks * sg.deg;
            }
            vn_path = std::max(vn_path, path);
            vn_work += path;
        }
        n_pos = (int)pbase;
        vn_idx.assign(ibase + 16, 0);

        // ---- check side ------------------------------------------------------------------------
        auto cs = schedule(cdeg, npw, warps, 1);
        edge_slot.assign(code.nnz, -1);
        size_t sbase = 0;
        ibase = 0;
        cn_path = cn_work = 0;
        for (int w = 0; w < warps; ++w)
        {
            encode(cs[w], csegs[w]);
            long path = 0;
            for (Seg &sg : csegs[w])
            {
                ibase = (ibase + 15) & ~(size_t)15;
                sg.base = (uint32_t)sbase;
                sg.idx = (uint32_t)ibase;
                sbase += (size_t)sg.ntasks * sg.deg * npw;
                ibase += (size_t)sg.ntasks * npw * idx_stride(sg.deg, isz);
                path += (long)sg.ntasks * sg.deg;
            }
            cn_path = std::max(cn_path, path);
            cn_work += path;
        }
        n_slots = (int)sbase;
        cn_idx.assign(ibase + 16, 0);
        if (n_slots >= (1 << 23) || n_pos >= (1 << 23)) throw std::runtime_error("layout too large");
        if (isz == 2 && ((size_t)n_slots * lanes > 65535 || (size_t)n_pos * lanes > 65535)) throw std::runtime_error("code too large for 16-bit indices");
        for (int w = 0; w < warps; ++w)
            for (const Seg &sg : csegs[w])
            {
                const int stride = idx_stride(sg.deg, isz);
                for (int t = 0; t < sg.ntasks; ++t)
                {
                    const Group &g = cs[w][sg.first + t];
                    for (int j = 0; j < (int)g.nodes.size(); ++j)
                    {
                        const int row = g.nodes[j];
                        for (int k = 0; k < sg.deg; ++k)
                        {
                            const int e = code.row_edge[code.row_ptr[row] + k];
                            edge_slot[e] = (int)(sg.base + ((size_t)t * sg.deg + k) * npw + j);
                            put(cn_idx, sg.idx + idx_offset(t, j, k, stride), var_pos[code.e_col[e]]);
                        }
                    }
                }
            }
        for (int w = 0; w < warps; ++w)
            for (const Seg &sg : vsegs[w])
            {
                const int stride = idx_stride(sg.deg, isz);
                for (int t = 0; t < sg.ntasks; ++t)
                {
                    const Group &g = vs[w][sg.first + t];
                    for (int j = 0; j < (int)g.nodes.size(); ++j)
                    {
                        const int col = g.nodes[j];
                        for (int k = 0; k < sg.deg; ++k)
                            put(vn_idx, sg.idx + idx_offset(t, j, k, stride), (uint32_t)edge_slot[code.col_edge[code.col_ptr[col] + k]]);
                    }
                }
            }
        flatten(csegs, cn_max_segs, cn_seg);
        flatten(vsegs, vn_max_segs, vn_seg);
    }
    // ------------------------------------------------------------------------------------------
    // BecSliceLayout: bank-conflict-free message slots for the bit-sliced erasure kernel.
    // The edges a warp touches in one shared-memory access are edge k of 32 consecutive checks (check phase) or of 32
    // consecutive variables (variable phase).  Make those access groups the nodes of a bipartite multigraph (left: check
    // groups, right: variable groups, one graph edge per code edge): every node has degree <= 32, so by Koenig's theorem the
    // edges have a proper 32-colouring.  With slot % 32 = colour, no access of either phase hits a bank twice.
    // ------------------------------------------------------------------------------------------
    void BecSliceLayout::build(const HostCode &code)
    {
        const int C = 32;
        const int md_c = std::max(code.max_cn_degree, 1), md_v = std::max(code.max_vn_degree, 1);
        const int nl = ((code.mc + 31) / 32) * md_c, nr = ((code.nc + 31) / 32) * md_v;
        std::vector<int> eu(code.nnz), ev(code.nnz); // graph endpoints of every code edge
        for (int c = 0; c < code.mc; ++c)
            for (int q = code.row_ptr[c]; q < code.row_ptr[c + 1]; ++q) eu[code.row_edge[q]] = (c / 32) * md_c + (q - code.row_ptr[c]);
        for (int v = 0; v < code.nc; ++v)
            for (int q = code.col_ptr[v]; q < code.col_ptr[v + 1]; ++q) ev[code.col_edge[q]] = (v / 32) * md_v + (q - code.col_ptr[v]);
        std::vector<int> L((size_t)nl * C, -1), R((size_t)nr * C, -1), colour(code.nnz, -1), used(C, 0);
        auto free_at = [&](const std::vector<int> &tab, int node) -> int
        { // the free colour of `node` whose class is smallest so far (keeps the classes, hence the padding, balanced)
            int best = -1;
            for (int k = 0; k < C; ++k)
                if (tab[(size_t)node * C + k] < 0 && (best < 0 || used[k] < used[best])) best = k;
            if (best < 0) throw std::runtime_error("edge colouring: node of degree > 32");
            return best;
        };
        std::vector<int> path;
        for (int e = 0; e < code.nnz; ++e)
        {
            const int u = eu[e], v = ev[e];
            const int a = free_at(L, u);
            int b = -1;
            if (R[(size_t)v * C + a] < 0) b = a;
            else b = free_at(R, v);
            if (a != b)
            { // colour a is taken at v: flip a <-> b along the alternating path that starts at v (it cannot reach u: it enters
              // left nodes through a-coloured edges and u has none)
                path.clear();
                int node = v, want = a;
                bool right = true;
                for (;;)
                {
                    const int f = right ? R[(size_t)node * C + want] : L[(size_t)node * C + want];
                    if (f < 0) break;
                    path.push_back(f);
                    node = right ? eu[f] : ev[f];
                    right = !right;
                    want = (want == a) ? b : a;
                }
                for (int f : path) { L[(size_t)eu[f] * C + colour[f]] = -1; R[(size_t)ev[f] * C + colour[f]] = -1; }
                for (int f : path)
                {
                    const int nc2 = (colour[f] == a) ? b : a;
                    --used[colour[f]]; ++used[nc2];
                    colour[f] = nc2;
                    L[(size_t)eu[f] * C + nc2] = f; R[(size_t)ev[f] * C + nc2] = f;
                }
            }
            colour[e] = a;
            ++used[a];
            L[(size_t)u * C + a] = e;
            R[(size_t)v * C + a] = e;
        }
        std::vector<int> next(C, 0), slot(code.nnz);
        for (int e = 0; e < code.nnz; ++e) slot[e] = C * next[colour[e]]++ + colour[e];
        n_slots = C * *std::max_element(next.begin(), next.end());
        if (n_slots > 65535) throw std::runtime_error("code too large for the bit-sliced erasure kernel");
        edge_slot = slot;
        row_slot.resize(code.nnz);
        col_slot.resize(code.nnz);
        for (int q = 0; q < code.nnz; ++q) { row_slot[q] = (uint16_t)slot[code.row_edge[q]]; col_slot[q] = (uint16_t)slot[code.col_edge[q]]; }
    }

    // ------------------------------------------------------------------------------------------
    // layered schedule
    // ------------------------------------------------------------------------------------------
    std::vector<std::vector<int>> auto_layers(const HostCode &code)
    {
        std::vector<std::vector<int>> layers;
        std::vector<std::vector<uint8_t>> occ; // occ[l][v]: variable v is used by a check of layer l
        for (int i = 0; i < code.mc; ++i)
        {
            size_t l = 0;
            for (;; ++l)
            {
                if (l == layers.size()) { layers.emplace_back(); occ.emplace_back(code.nc, 0); }
                bool clash = false;
                for (int q = code.row_ptr[i]; q < code.row_ptr[i + 1] && !clash; ++q) clash = occ[l][code.e_col[code.row_edge[q]]] != 0;
                if (!clash) break;
            }
            for (int q = code.row_ptr[i]; q < code.row_ptr[i + 1]; ++q) occ[l][code.e_col[code.row_edge[q]]] = 1;
            layers[l].push_back(i);
        }
        return layers;
    }

    std::vector<std::vector<int>> read_layer_file(const std::string &path)
    {
        std::ifstream f(path);
        if (!f.good()) throw std::runtime_error("can not open layer file for reading");
        std::string tok;
        auto after_colon = [&](const char *what) -> long
        { // "<key>: <value>"
            std::string line;
            while (std::getline(f, line))
            {
                const size_t c = line.find(':');
                if (c == std::string::npos) continue;
                return std::stol(line.substr(c + 1));
            }
            throw std::runtime_error(std::string("layer file: missing ") + what);
        };
        const long nl = after_colon("nl");
        if (nl < 1 || nl > (1 << 20)) throw std::runtime_error("layer file: bad layer count");
        std::vector<std::vector<int>> layers((size_t)nl);
        for (long l = 0; l < nl; ++l)
        {
            const long w = after_colon("cn[i]");
            if (w < 0) throw std::runtime_error("layer file: bad layer size");
            layers[l].resize((size_t)w);
            for (long k = 0; k < w; ++k)
                if (!(f >> layers[l][k])) throw std::runtime_error("layer file: truncated layer");
            std::string rest;
            std::getline(f, rest);
        }
        return layers;
    }

    void validate_layers(const HostCode &code, const std::vector<std::vector<int>> &layers)
    {
        std::vector<int> seen(code.nc, -1), used(code.mc, 0);
        for (size_t l = 0; l < layers.size(); ++l)
            for (int chk : layers[l])
            {
                if (chk < 0 || chk >= code.mc) throw std::runtime_error("layers: check index out of range");
                if (used[chk]++) throw std::runtime_error("layers: a check appears twice");
                for (int q = code.row_ptr[chk]; q < code.row_ptr[chk + 1]; ++q)
                {
                    const int v = code.e_col[code.row_edge[q]];
                    if (seen[v] == (int)l)
                        throw std::runtime_error("layers: two checks of one layer share a variable (the in-place layered update needs disjoint checks per layer)");
                    seen[v] = (int)l;
                }
            }
        for (int i = 0; i < code.mc; ++i)
            if (!used[i]) throw std::runtime_error("layers: a check belongs to no layer");
    }

    void LayeredLayout::build(const HostCode &code, const std::vector<std::vector<int>> &layers, int lanes_, int threads_)
    {
        if (lanes_ != 1 && lanes_ != 2 && lanes_ != 4) throw std::runtime_error("lanes per node must be 1, 2 or 4");
        if (threads_ < 32 || threads_ > 512 || threads_ % 32) throw std::runtime_error("threads_per_cta must be a multiple of 32 <= 512");
        if (code.max_cn_degree > 64) throw std::runtime_error("check degree > 64 not supported");
        if (code.min_cn_degree < 2) throw std::runtime_error("check nodes of degree < 2 are not supported (undefined in the reference)");
        validate_layers(code, layers);
        lanes = lanes_; threads = threads_; warps = threads / 32; npw = 32 / lanes; n_layers = (int)layers.size();
        struct Seg { int deg, cnt, ntasks; uint32_t base; };
        std::vector<std::vector<Seg>> lists((size_t)n_layers * warps);
        std::vector<std::vector<std::vector<int>>> nodes((size_t)n_layers * warps); // per list, per task: the checks
        edge_slot.assign(code.nnz, -1);
        uint32_t next = 0;
        max_segs = 1;
        for (int l = 0; l < n_layers; ++l)
        {
            std::vector<int> order(layers[l]);
            std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
                return code.row_ptr[a + 1] - code.row_ptr[a] > code.row_ptr[b + 1] - code.row_ptr[b]; });
            // tasks of npw equal-degree checks, dealt to the least loaded warp (load in edges)
            std::vector<long> load(warps, 0);
            for (size_t i = 0; i < order.size();)
            {
                const int deg = code.row_ptr[order[i] + 1] - code.row_ptr[order[i]];
                std::vector<int> task;
                while (i < order.size() && code.row_ptr[order[i] + 1] - code.row_ptr[order[i]] == deg && (int)task.size() < npw) task.push_back(order[i++]);
                const int w = (int)(std::min_element(load.begin(), load.end()) - load.begin());
                load[w] += deg;
                auto &lst = lists[(size_t)l * warps + w];
                auto &nds = nodes[(size_t)l * warps + w];
                if (!lst.empty() && lst.back().deg == deg && lst.back().cnt == (int)task.size() && lst.back().cnt == npw && lst.back().ntasks < 65535) ++lst.back().ntasks;
                else lst.push_back({deg, (int)task.size(), 1, 0u});
                nds.push_back(std::move(task));
            }
            for (int w = 0; w < warps; ++w)
            {
                auto &lst = lists[(size_t)l * warps + w];
                auto &nds = nodes[(size_t)l * warps + w];
                max_segs = std::max<int>(max_segs, (int)lst.size() + 1);
                size_t t0 = 0;
                for (Seg &sg : lst)
                {
                    sg.base = next;
                    for (int t = 0; t < sg.ntasks; ++t)
                        for (int j = 0; j < (int)nds[t0 + t].size(); ++j)
                        {
                            const int row = nds[t0 + t][j];
                            for (int k = 0; k < sg.deg; ++k)
                                edge_slot[code.row_edge[code.row_ptr[row] + k]] = (int)(sg.base + ((uint32_t)t * sg.deg + k) * npw + j);
                        }
                    next += (uint32_t)sg.ntasks * sg.deg * npw;
                    t0 += sg.ntasks;
                }
            }
        }
        n_slots = (int)next;
        if ((size_t)n_slots * 16 * lanes >= (1ull << 32) || (size_t)code.nc * 16 * lanes >= (1ull << 32)) throw std::runtime_error("layout too large");
        idx.assign(std::max(n_slots, 1), 0);
        for (int e = 0; e < code.nnz; ++e) idx[edge_slot[e]] = (uint32_t)code.e_col[e] * (uint32_t)(16 * lanes);
        seg.assign((size_t)4 * n_layers * warps * max_segs, 0);
        for (size_t li = 0; li < lists.size(); ++li)
            for (size_t sgi = 0; sgi < lists[li].size(); ++sgi)
            {
                const Seg &g = lists[li][sgi];
                uint32_t *d = &seg[4 * (li * max_segs + sgi)];
                d[0] = (uint32_t)g.deg | ((uint32_t)g.cnt << 8) | ((uint32_t)g.ntasks << 16);
                d[1] = g.base;
            }
    }
} // namespace b200
