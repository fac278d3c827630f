/*
 * flat_oracle.c — CPU restatement of the FlatDD array-phase algorithms on flat DD tables.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under flatdd_b200/ links, imports or calls this file;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg do, and there only as
 * the checker.  Parity status: PINNED — every function here is checked bit-for-bit against
 * outputs of the compiled reference itself (oracle/_ref, built from /root/reference by
 * oracle/Makefile; fixtures under tests/golden/ written by oracle/ref_dump.cpp through
 * oracle/make_golden.py).  The reference ships no tests or golden vectors of its own
 * (SURVEY.md section 4), so the compiled reference is the only pin there is.
 *
 * Each function cites the reference file:line it follows (paths relative to the reference
 * tree).  Arithmetic is written as separate multiplies and adds in the reference's order;
 * compile with -ffp-contract=off so the compiler cannot fuse them (the reference is built
 * with -mavx and no FMA, CMakeLists.txt:20).
 */
#include "flatdd_b200.h"

#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define API __attribute__((visibility("default")))

static int is_zero_w(const double* w) { return w[0] == 0.0 && w[1] == 0.0; }

/* --------------------------------------------------------------------------------------------
 * DD -> array, exact per-amplitude walk.
 * Follows getValueByPathPar (include/dd/SwitchPackage.hpp:3605-3634): c = 1; at every node
 * c = c * w(incoming edge), then descend along bit v of the index; after the terminal edge is
 * reached multiply its weight too.  The destination index i runs 0..2^n-1 with qubit q = bit q
 * (include/SwitchSimulator.hpp:72-80, 305-311).  Like the reference's parallel conversion we
 * leave amplitudes of zero sub-trees at +0.0 (it never writes them, :269-278).
 * ------------------------------------------------------------------------------------------ */
API int oracle_convert(const fdd_vecdd* dd, double* re, double* im) {
    const int n = dd->n_qubits;
    const uint64_t dim = 1ULL << n;
    for (uint64_t i = 0; i < dim; ++i) {
        double cr = 1.0, ci = 0.0;
        const double* w = dd->root_weight;
        int32_t node = dd->root;
        int dead = 0;
        for (;;) {
            /* c = c * w  (SwitchPackage.hpp:3616-3621) */
            const double dr = cr * w[0] - ci * w[1];
            const double di = cr * w[1] + ci * w[0];
            cr = dr;
            ci = di;
            if (is_zero_w(w)) { dead = 1; break; }
            if (node == FDD_TERMINAL) break;
            const int bit = (int)((i >> dd->level[node]) & 1ULL);
            w = dd->weight + 2 * (2 * (size_t)node + bit);
            node = dd->child[2 * (size_t)node + bit];
        }
        re[i] = dead ? 0.0 : cr;
        im[i] = dead ? 0.0 : ci;
    }
    return 0;
}

/* --------------------------------------------------------------------------------------------
 * DD -> array with the reference's "initial regularity" shortcut.
 * Follows getVectorFromDDSwitch1 (include/SwitchSimulator.hpp:169-352): while the current
 * node's two successors are the same node, the index range is split into segments with an
 * accumulated weight (:208-250, stops once more than 2^n_thread_exp segments exist); only
 * the first segment is walked (:252-262, 293-317), every other segment is the first one
 * times w_seg / w_first (:319-351; the AVX loop drops a tail of < 4 elements, `& ~0x3`).
 * The BFS that distributes sub-trees over threads (:264-290) does not change any value, so
 * it is not restated.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int32_t node; /* edge.p */
    const double* ew; /* edge.w */
    double real, imag;
    uint64_t beg, end;
} seg_t;

API int oracle_convert_switch1(const fdd_vecdd* dd, unsigned n_thread_exp, double* re, double* im) {
    const int n = dd->n_qubits;
    const uint64_t dim = 1ULL << n;
    const size_t n_thread = (size_t)1 << n_thread_exp;
    memset(re, 0, sizeof(double) * dim);
    memset(im, 0, sizeof(double) * dim);
    oracle_convert(dd, re, im); /* every walked amplitude has exactly this value */

    /* segments, kept in the lexicographic order of their bit-string keys like the std::map */
    size_t cap = 2 * n_thread + 4, n_last = 1, n_this = 0;
    seg_t* last = (seg_t*)malloc(sizeof(seg_t) * cap);
    seg_t* cur = (seg_t*)malloc(sizeof(seg_t) * cap);
    last[0].node = dd->root;
    last[0].ew = dd->root_weight;
    last[0].real = dd->root_weight[0];
    last[0].imag = dd->root_weight[1];
    last[0].beg = 0;
    last[0].end = dim;
    int seg_exp = n;
    int ir_flag = 0;
    int32_t this_node = dd->root;
    while (this_node != FDD_TERMINAL && dd->child[2 * (size_t)this_node] == dd->child[2 * (size_t)this_node + 1]) {
        ir_flag = 1;
        n_this = 0;
        for (size_t s = 0; s < n_last; ++s) {
            const seg_t* e = &last[s];
            for (int b = 0; b < 2; ++b) {
                const double* w = dd->weight + 2 * (2 * (size_t)e->node + b);
                if (is_zero_w(w)) continue;
                seg_t* t = &cur[n_this++];
                t->node = dd->child[2 * (size_t)e->node + b];
                t->ew = w;
                t->real = e->real * w[0] - e->imag * w[1];
                t->imag = e->real * w[1] + e->imag * w[0];
                const uint64_t half = 1ULL << (seg_exp - 1);
                t->beg = b ? e->beg + half : e->beg;
                t->end = b ? e->end : e->beg + half;
                this_node = t->node;
            }
        }
        if (n_this > n_thread) break;
        seg_exp--;
        seg_t* tmp = last; last = cur; cur = tmp;
        n_last = n_this;
        if (this_node == FDD_TERMINAL) break; /* reference would dereference nullptr here; n==depth guard */
    }
    if (ir_flag && n_last > 1) {
        const seg_t first = last[0];
        for (size_t s = 1; s < n_last; ++s) {
            const seg_t* e = &last[s];
            const double den = first.real * first.real + first.imag * first.imag;
            const double common_r = (e->real * first.real + e->imag * first.imag) / den;
            const double common_i = (e->imag * first.real - e->real * first.imag) / den;
            const uint64_t len = (e->end - e->beg) & ~(uint64_t)0x3;
            for (uint64_t k = 0; k < len; ++k) {
                const double zr = re[first.beg + k], zi = im[first.beg + k];
                /* resr = zr*cr + (-(zi*ci)); resi = zr*ci + zi*cr  (:334-340) */
                re[e->beg + k] = zr * common_r + (-(zi * common_i));
                im[e->beg + k] = zr * common_i + zi * common_r;
            }
            /* the dropped tail keeps the pre-zeroed value */
            for (uint64_t k = len; k < e->end - e->beg; ++k) {
                re[e->beg + k] = 0.0;
                im[e->beg + k] = 0.0;
            }
        }
    }
    free(last);
    free(cur);
    return 0;
}

/* --------------------------------------------------------------------------------------------
 * DMAVM, z += M * y.
 * Follows DDArrMultiplyIP -> DDArrMultiplyIP2 -> AssignParVectorIP -> DDArrMultiplyRecurIP
 * (include/dd/SwitchPackage.hpp:1897-1925, 2132-2204, 2207-2261).  AssignParVectorIP starts
 * with fact = (1, 0) and multiplies the weights of the top levels left to right (:2187-2193);
 * DDArrMultiplyRecurIP continues the same product (:2242-2248), forms the last factor at the
 * terminal (:2221-2227), multiplies by y (:2229-2233) and accumulates (:2235-2236).  Sub-blocks
 * are visited row half first, column half second, so every z[r] is summed over ascending
 * columns; the thread split only decides which thread runs which rows, hence the result is
 * independent of the thread count and this single-threaded recursion is bit-identical.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    const fdd_matdd* g;
    const double *yr, *yi;
    double *zr, *zi;
} mv_t;

static void dmavm_rec(const mv_t* c, int32_t node, const double* w, int var, uint64_t col0, uint64_t row0,
                      double fr, double fi) {
    if (is_zero_w(w)) return; /* :2214 */
    const double z1r = fr * w[0] - fi * w[1];
    const double z1i = fr * w[1] + fi * w[0];
    if (node == FDD_TERMINAL) { /* :2218-2238 */
        const double y_r = c->yr[col0], y_i = c->yi[col0];
        const double z2r = z1r * y_r - z1i * y_i;
        const double z2i = z1r * y_i + z1i * y_r;
        c->zr[row0] += z2r;
        c->zi[row0] += z2i;
        return;
    }
    const uint64_t half = 1ULL << var; /* var == level of `node` (:2251) */
    for (int i = 0; i < 2; ++i) {    /* :2249-2260 */
        for (int j = 0; j < 2; ++j) {
            const size_t e = 4 * (size_t)node + 2 * i + j;
            dmavm_rec(c, c->g->child[e], c->g->weight + 2 * e, var - 1, col0 + j * half, row0 + i * half, z1r, z1i);
        }
    }
}

/* z must be zeroed by the caller when `accumulate` != 0 (the reference contract); with
 * accumulate == 0 it is zeroed here. */
API int oracle_dmavm(const fdd_matdd* g, const double* y_real, const double* y_imag, double* z_real,
                     double* z_imag, int accumulate) {
    const uint64_t dim = 1ULL << g->n_qubits;
    if (!accumulate) {
        memset(z_real, 0, sizeof(double) * dim);
        memset(z_imag, 0, sizeof(double) * dim);
    }
    mv_t c = {g, y_real, y_imag, z_real, z_imag};
    dmavm_rec(&c, g->root, g->root_weight, g->n_qubits - 1, 0, 0, 1.0, 0.0);
    return 0;
}

/* --------------------------------------------------------------------------------------------
 * Cost model.
 * oracle_mac_count: DMAVMACCountIP (include/dd/SwitchPackage.hpp:3285-3311) — memoised number
 * of non-zero root-to-terminal paths; a terminal counts 1.
 * ------------------------------------------------------------------------------------------ */
static uint64_t mac_rec(const fdd_matdd* g, int32_t node, uint64_t* memo, uint8_t* have) {
    if (node == FDD_TERMINAL) return 1;
    if (have[node]) return memo[node];
    uint64_t cnt = 0;
    for (int i = 0; i < 2; ++i) {
        for (int j = 0; j < 2; ++j) {
            const size_t e = 4 * (size_t)node + 2 * i + j;
            if (!is_zero_w(g->weight + 2 * e)) cnt += mac_rec(g, g->child[e], memo, have);
        }
    }
    memo[node] = cnt;
    have[node] = 1;
    return cnt;
}

API uint64_t oracle_mac_count(const fdd_matdd* g) {
    uint64_t* memo = (uint64_t*)calloc((size_t)g->n_nodes + 1, sizeof(uint64_t));
    uint8_t* have = (uint8_t*)calloc((size_t)g->n_nodes + 1, 1);
    const uint64_t r = mac_rec(g, g->root, memo, have);
    free(memo);
    free(have);
    return r;
}

/* DMAVMACStatIP (:3012-3017): nnz / 2^n_thread_exp (integer division). */
API uint64_t oracle_cost_ip(const fdd_matdd* g, unsigned n_thread_exp) {
    return oracle_mac_count(g) / ((uint64_t)1 << n_thread_exp);
}

/* DMAVMACStatOP1 -> DMAVMACCountOP1 (:3006-3010, 3202-3283).
 * AssignParVectorOP (:2472-2508) splits the top n_thread_exp levels by COLUMN bit (outer loop
 * i) and row bit (inner loop j, edge e[j*2+i]); each surviving edge is appended to the list of
 * its column block together with its output range [row_off, row_off + seg).  The greedy
 * first-fit (:3219-3262) then packs column blocks into scratch buffers whose claimed ranges do
 * not overlap; note that it iterates over BOTH stored offsets (start and end) of every edge
 * as if each were a start, and stores them as int — restated as is.  Cost (:3266-3282): per
 * column block, a sub-DD seen for the first time costs its nnz, a repeated one costs seg/4;
 * the sum is divided by the thread count and nDim*num_buf/(4*threads) is added for the merge. */
typedef struct { int32_t node; uint64_t off; } opent_t;
typedef struct { opent_t* v; size_t n, cap; } oplist_t;

static void op_push(oplist_t* l, int32_t node, uint64_t off) {
    if (l->n == l->cap) {
        l->cap = l->cap ? 2 * l->cap : 8;
        l->v = (opent_t*)realloc(l->v, sizeof(opent_t) * l->cap);
    }
    l->v[l->n].node = node;
    l->v[l->n].off = off;
    l->n++;
}

static void assign_op(const fdd_matdd* g, int32_t node, const double* w, oplist_t* lists, unsigned max_lev,
                      unsigned cur_lev, size_t col_off, int var, uint64_t row_off) {
    if (is_zero_w(w)) return;
    if (max_lev == cur_lev) {
        op_push(&lists[col_off], node, row_off);
        return;
    }
    for (size_t i = 0; i < 2; ++i) {
        for (size_t j = 0; j < 2; ++j) {
            const size_t e = 4 * (size_t)node + j * 2 + i;
            assign_op(g, g->child[e], g->weight + 2 * e, lists, max_lev, cur_lev + 1,
                      col_off + i * ((size_t)1 << (max_lev - cur_lev - 1)), var,
                      row_off + ((uint64_t)1 << (var - (int)cur_lev)) * j);
        }
    }
}

typedef struct { int first, second; } ipair_t;
typedef struct { ipair_t* v; size_t n, cap; } ipairs_t;

static void ip_push(ipairs_t* l, int a, int b) {
    if (l->n == l->cap) {
        l->cap = l->cap ? 2 * l->cap : 8;
        l->v = (ipair_t*)realloc(l->v, sizeof(ipair_t) * l->cap);
    }
    l->v[l->n].first = a;
    l->v[l->n].second = b;
    l->n++;
}

API uint64_t oracle_cost_op1(const fdd_matdd* g, unsigned n_thread_exp) {
    const uint64_t n_dim = 1ULL << g->n_qubits;
    const size_t n_thread = (size_t)1 << n_thread_exp;
    const uint64_t seg = n_dim / n_thread;
    oplist_t* lists = (oplist_t*)calloc(n_thread, sizeof(oplist_t));
    assign_op(g, g->root, g->root_weight, lists, n_thread_exp, 0, 0, g->n_qubits - 1, 0);

    /* greedy scratch-buffer packing */
    ipairs_t* buffers = NULL;
    size_t n_buf_alloc = 0;
    int num_buf = 0;
    for (size_t i = 0; i < n_thread; ++i) {
        /* yoffset_vec[i] = {off, off+seg, off, off+seg, ...} */
        const size_t n_off = 2 * lists[i].n;
        int add_to = -1;
        for (size_t j = 0; j < (size_t)num_buf && add_to < 0; ++j) {
            int can_add = 1;
            for (size_t k = 0; k < buffers[j].n && can_add; ++k) {
                for (size_t t = 0; t < n_off; ++t) {
                    const uint64_t yo = lists[i].v[t / 2].off + ((t & 1) ? seg : 0);
                    /* comparison happens in size_t after converting the int pair members */
                    if ((uint64_t)(int64_t)buffers[j].v[k].first < yo + seg &&
                        (uint64_t)(int64_t)buffers[j].v[k].second > yo) {
                        can_add = 0;
                        break;
                    }
                }
            }
            if (can_add) {
                add_to = (int)j;
                for (size_t t = 0; t < n_off; ++t) {
                    const uint64_t yo = lists[i].v[t / 2].off + ((t & 1) ? seg : 0);
                    ip_push(&buffers[j], (int)yo, (int)(yo + seg));
                }
                /* std::sort by .first: order does not influence the overlap test */
            }
        }
        if (add_to == -1) {
            if ((size_t)num_buf == n_buf_alloc) {
                n_buf_alloc = n_buf_alloc ? 2 * n_buf_alloc : 4;
                buffers = (ipairs_t*)realloc(buffers, sizeof(ipairs_t) * n_buf_alloc);
            }
            buffers[num_buf].v = NULL;
            buffers[num_buf].n = buffers[num_buf].cap = 0;
            for (size_t t = 0; t < n_off; ++t) {
                const uint64_t yo = lists[i].v[t / 2].off + ((t & 1) ? seg : 0);
                ip_push(&buffers[num_buf], (int)yo, (int)(yo + seg));
            }
            num_buf++;
        }
    }

    uint64_t* memo = (uint64_t*)calloc((size_t)g->n_nodes + 1, sizeof(uint64_t));
    uint8_t* have = (uint8_t*)calloc((size_t)g->n_nodes + 1, 1);
    uint8_t* visited = (uint8_t*)malloc((size_t)g->n_nodes + 1);
    uint64_t cnt = 0;
    for (size_t i = 0; i < n_thread; ++i) {
        memset(visited, 0, (size_t)g->n_nodes + 1);
        for (size_t j = 0; j < lists[i].n; ++j) {
            const int32_t node = lists[i].v[j].node;
            const size_t slot = node == FDD_TERMINAL ? (size_t)g->n_nodes : (size_t)node;
            if (visited[slot]) {
                cnt += seg / 4;
            } else {
                visited[slot] = 1;
                /* mac_map[p]: the memoised count; a terminal is never inserted, so it reads 0 */
                cnt += node == FDD_TERMINAL ? 0 : mac_rec(g, node, memo, have);
            }
        }
    }
    const uint64_t result = cnt / n_thread + n_dim * (uint64_t)num_buf / (4 * n_thread);
    for (int b = 0; b < num_buf; ++b) free(buffers[b].v);
    free(buffers);
    for (size_t i = 0; i < n_thread; ++i) free(lists[i].v);
    free(lists);
    free(memo);
    free(have);
    free(visited);
    return result;
}

/* --------------------------------------------------------------------------------------------
 * DD size and the switch rule.
 * oracle_dd_size: size() / nodeCount (include/dd/SwitchPackage.hpp:2992-2998, 3156-3169) —
 * distinct nodes reachable from the root, the terminal counted once (every DD of depth >= 1
 * reaches it).  Works on either table via the stride (2 or 4).
 * oracle_switch_index: the EMA rule of singleShot (src/SwitchSimulator.cpp:97, 163-165, 181):
 * EMA_0 = n; after measurement k of the DD size s_k: switch iff EMA > 0 && EMA*threshold < s_k
 * (tested with the OLD EMA), then EMA = beta*EMA + (1-beta)*s_k.  Returns the first k that
 * switches or -1.
 * ------------------------------------------------------------------------------------------ */
API uint64_t oracle_dd_size(int32_t n_nodes, int32_t root, const int32_t* child, int radix) {
    if (root == FDD_TERMINAL) return 1;
    uint8_t* seen = (uint8_t*)calloc((size_t)n_nodes, 1);
    int32_t* stack = (int32_t*)malloc(sizeof(int32_t) * ((size_t)n_nodes * radix + 1));
    size_t sp = 0;
    uint64_t count = 1; /* the terminal */
    stack[sp++] = root;
    seen[root] = 1;
    while (sp) {
        const int32_t u = stack[--sp];
        count++;
        for (int k = 0; k < radix; ++k) {
            const int32_t c = child[(size_t)u * radix + k];
            if (c != FDD_TERMINAL && !seen[c]) {
                seen[c] = 1;
                stack[sp++] = c;
            }
        }
    }
    free(seen);
    free(stack);
    return count;
}

API int oracle_switch_index(int n_qubits, double beta, double threshold, const double* sizes, int n_sizes) {
    double ema = (double)n_qubits;
    for (int k = 0; k < n_sizes; ++k) {
        const double s = sizes[k];
        const double new_v = ema * beta + (1 - beta) * s;
        if (ema > 0 && ema * threshold < s) return k;
        ema = new_v;
    }
    return -1;
}
