// kernels.cuh -- device code of libmsb200 (sm_100a).
//
//   encode_pack_kernel      ASCII -> 2-bit codes + N mask            (cscore.c:81-114)
//   prefilter_kernel        every window x every motif x both strands, 16-bit fixed point,
//                           conservative: emits a superset of the hits  (cscore.c:336-355)
//   exact_candidates_kernel fp64 re-score of candidates in the reference's accumulation
//                           order + the reference's hit predicate       (cscore.c:344-389)
//   exact_dirty_kernel      the same for windows that touch a non-ACGT base
//   exact_slow_kernel       the same for motifs the table path cannot take (L > 32, non-finite)
//   decode_sites_kernel     sorted keys -> (seq_idx, start, strand); motif_offsets_kernel: CSR
//   dedup_flags_kernel      adjacent-site de-duplication              (scanner.py:156-193)
//   score0_kernel           offset-0 window score of every sequence     (cscore.c:191-223)
#pragma once
#include "common.cuh"

namespace msb {

// Largest s with poff[s] <= p.  Empty sequences share a start with their successor and are
// skipped because the search returns the LAST such s.
__device__ __forceinline__ int64_t find_seq(const int64_t *__restrict__ poff, int64_t n_seqs,
                                            int64_t p) {
    int64_t lo = 0, hi = n_seqs;  // invariant: poff[lo] <= p, answer in [lo, hi)
    while (hi - lo > 1) {
        int64_t mid = (lo + hi) >> 1;
        if (__ldg(poff + mid) <= p) lo = mid; else hi = mid;
    }
    return lo;
}

// The reference's int8 code of packed position q: 0..3, or -1 for anything but ACGT/acgt.
__device__ __forceinline__ int base_code(const SeqView &S, int64_t q) {
    uint32_t m = __ldg(S.nmask + (q >> 5));
    if ((m >> (q & 31)) & 1u) return -1;
    uint32_t w = __ldg(S.codes + (q >> 4));
    return (int) ((w >> ((q & 15) * 2)) & 3u);
}

// ---------------------------------------------------------------------------------------------
// encode + pack.  One thread per 32-base block of the packed space.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
encode_pack_kernel(const uint8_t *__restrict__ ascii, const int64_t *__restrict__ seq_off,
                   const int64_t *__restrict__ poff, const int32_t *__restrict__ len,
                   int64_t n_seqs, int64_t first_block, int64_t n_blocks, uint32_t *__restrict__ codes,
                   uint32_t *__restrict__ nmask, int32_t *__restrict__ blk_seq) {
    int64_t b = first_block + (int64_t) blockIdx.x * blockDim.x + threadIdx.x;   // blocks [first_block, n_blocks)
    if (b >= n_blocks) return;
    int64_t p = b * kPadBases;
    int64_t s = find_seq(poff, n_seqs, p);
    int64_t j = p - __ldg(poff + s);
    int n = (int) min((int64_t) kPadBases, (int64_t) __ldg(len + s) - j);
    const uint8_t *src = ascii + __ldg(seq_off + s) + j;
    uint32_t lo = 0, hi = 0, mask = 0;
#pragma unroll
    for (int k = 0; k < kPadBases; k++) {
        if (k < n) {
            uint32_t c = (uint32_t) __ldg(src + k) | 0x20u;  // fold case (cscore.c:91-108)
            uint32_t code = 0, isn = 0;
            if (c == 0x61u) code = 0;        // A a
            else if (c == 0x63u) code = 1;   // C c
            else if (c == 0x67u) code = 2;   // G g
            else if (c == 0x74u) code = 3;   // T t
            else isn = 1;                    // default: -1 (cscore.c:109-110)
            if (k < 16) lo |= code << (2 * k); else hi |= code << (2 * (k - 16));
            mask |= isn << k;
        }
    }
    codes[2 * b] = lo;
    codes[2 * b + 1] = hi;
    nmask[b] = mask;
    blk_seq[b] = (int32_t) s;
}

// ---------------------------------------------------------------------------------------------
// extract: build a packed sequence set from intervals of a resident one (the device-side
// counterpart of Scanner._extract_seq's per-region fetch, scanner.py:71-87).  One thread per
// 32-base block of the destination: 64 code bits + 32 mask bits funnel-shifted out of the source.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
extract_pack_kernel(SeqView src, const int32_t *__restrict__ src_idx, const int64_t *__restrict__ src_start,
                    const int64_t *__restrict__ poff, const int32_t *__restrict__ len, int64_t n_seqs,
                    int64_t n_blocks, uint32_t *__restrict__ codes, uint32_t *__restrict__ nmask,
                    int32_t *__restrict__ blk_seq) {
    const int64_t b = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blocks) return;
    const int64_t p = b * kPadBases;
    const int64_t s = find_seq(poff, n_seqs, p);
    const int64_t j = p - __ldg(poff + s);
    const int n = (int) min((int64_t) kPadBases, (int64_t) __ldg(len + s) - j);
    uint32_t lo = 0, hi = 0, mask = 0;
    if (n > 0) {
        const int64_t q = __ldg(src.poff + __ldg(src_idx + s)) + __ldg(src_start + s) + j;
        // the source buffers carry zero padding behind their last block (msb_seqs_from_ascii)
        const uint32_t *cp = src.codes + (q >> 4);
        const uint32_t sh = (uint32_t) (q & 15) * 2;
        const uint32_t c0 = __ldg(cp), c1 = __ldg(cp + 1), c2 = __ldg(cp + 2);
        lo = __funnelshift_r(c0, c1, sh);
        hi = __funnelshift_r(c1, c2, sh);
        const uint32_t *mp = src.nmask + (q >> 5);
        mask = __funnelshift_r(__ldg(mp), __ldg(mp + 1), (uint32_t) (q & 31));
        if (n < 32) {
            mask &= (1u << n) - 1u;
            if (n <= 16) { hi = 0; lo &= n == 16 ? 0xffffffffu : ((1u << (2 * n)) - 1u); }
            else hi &= (1u << (2 * (n - 16))) - 1u;
        }
    }
    codes[2 * b] = lo;
    codes[2 * b + 1] = hi;
    nmask[b] = mask;
    blk_seq[b] = (int32_t) s;
}

// Sequences uploaded in the packed representation (msb_seqs_from_packed) -> canonical form: bits behind a
// sequence's last base cleared in both planes, codes of non-ACGT bases cleared (N -> 0), block owners filled in.
__device__ __forceinline__ uint32_t spread_bits16(uint32_t x) {   // bit i of the low half -> bit 2 i
    x = (x | (x << 8)) & 0x00FF00FFu;
    x = (x | (x << 4)) & 0x0F0F0F0Fu;
    x = (x | (x << 2)) & 0x33333333u;
    x = (x | (x << 1)) & 0x55555555u;
    return x;
}
__global__ void __launch_bounds__(256)
packed_fixup_kernel(const int64_t *__restrict__ poff, const int32_t *__restrict__ len, int64_t n_seqs,
                    int64_t n_blocks, uint32_t *__restrict__ codes, uint32_t *__restrict__ nmask,
                    int32_t *__restrict__ blk_seq) {
    const int64_t b = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blocks) return;
    const int64_t p = b * kPadBases;
    const int64_t s = find_seq(poff, n_seqs, p);
    const int64_t left = (int64_t) __ldg(len + s) - (p - __ldg(poff + s));
    const int n = left >= kPadBases ? kPadBases : (left > 0 ? (int) left : 0);
    const uint32_t valid = n >= 32 ? 0xffffffffu : ((1u << n) - 1u);
    const uint32_t mask = nmask[b] & valid;
    const uint32_t acgt = valid & ~mask;   // bases whose code counts
    codes[2 * b] &= 3u * spread_bits16(acgt & 0xffffu);
    codes[2 * b + 1] &= 3u * spread_bits16(acgt >> 16);
    nmask[b] = mask;
    blk_seq[b] = (int32_t) s;
}

// Number of non-ACGT bases in each of n windows [start, start + length) of a resident set (the
// acceptance test of the background sampler, genome/__init__.py:172-175, counts the N of a sample).
// One thread per window; the window is clipped at its sequence's end.
__global__ void __launch_bounds__(256)
window_ncount_kernel(SeqView src, const int32_t *__restrict__ src_idx, const int64_t *__restrict__ start,
                     int32_t length, int64_t n, int32_t *__restrict__ counts) {
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t s = __ldg(src_idx + i), a = __ldg(start + i);
    const int64_t left = (int64_t) __ldg(src.len + s) - a;
    int64_t todo = left < length ? left : length;
    int64_t q = __ldg(src.poff + s) + a;
    int32_t c = 0;
    while (todo > 0) {
        const int off = (int) (q & 31);
        const int take = (int) (todo < 32 - off ? todo : 32 - off);
        const uint32_t w = __ldg(src.nmask + (q >> 5)) >> off;
        c += __popc(take >= 32 ? w : (w & ((1u << take) - 1u)));
        q += take;
        todo -= take;
    }
    counts[i] = c;
}

// Parity accessor: packed -> int8 codes laid out like the ASCII input.
__global__ void __launch_bounds__(256)
unpack_codes_kernel(SeqView S, const int64_t *__restrict__ seq_off, int8_t *__restrict__ out) {
    int64_t p = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= S.total_packed) return;
    int64_t s = find_seq(S.poff, S.n_seqs, p);
    int64_t j = p - S.poff[s];
    if (j < S.len[s]) out[seq_off[s] + j] = (int8_t) base_code(S, p);
}

// ---------------------------------------------------------------------------------------------
// prefilter.
//
// Work unit: a warp chunk of 32*W consecutive packed positions; lane l owns the W window
// starts p0 .. p0+W-1 with p0 = chunk*32*W + l*W.  A thread keeps the pre-scaled 2-mer index of
// every offset it can need (W + 2*(kMaxGroups-1) registers), so the per-motif inner loop is
// one LDS [idx + motif_base + imm] plus a third of an IADD3 per 2 PWM columns of BOTH strands:
// a table entry packs (forward sum << 16) + reverse sum of its two columns in 16-bit fixed
// point, pre-biased so that bit 31 / bit 15 of the final sum is the "forward / reverse score
// may reach the cutoff" flag.  Entries of one group are 16 consecutive words = 16 banks, lanes
// that read the same word are broadcast: conflict-free by construction.
// ---------------------------------------------------------------------------------------------
template <int W> struct PF {
    static constexpr int kIdx = W + 2 * (kMaxGroups - 1);  // 2-mer offsets a thread may use
    static constexpr int kThreads = 512;
    static constexpr int kChunk = 32 * W;
};

template <int G, int W>
__device__ __forceinline__ uint32_t score_motif(const unsigned char *__restrict__ tab,
                                                const uint32_t (&idx)[PF<W>::kIdx],
                                                uint32_t (&acc)[W]) {
    uint32_t any = 0;
#pragma unroll
    for (int w = 0; w < W; w++) {
        uint32_t a = 0;
#pragma unroll
        for (int g = 0; g < G; g++)
            a += *reinterpret_cast<const uint32_t *>(tab + g * kGroupBytes + idx[w + 2 * g]);
        acc[w] = a;
        any |= a;
    }
    return any & 0x80008000u;
}

template <int W>
__device__ __forceinline__ void emit_candidates(const PrefilterParams &P,
                                                const uint32_t (&acc)[W], uint32_t live,
                                                uint32_t sorted_idx, int64_t p0, int32_t j0,
                                                int32_t slen) {
    int32_t L = __ldg(P.mlen + __ldg(P.order + sorted_idx));
#pragma unroll
    for (int w = 0; w < W; w++) {
        uint32_t f = acc[w] & 0x80008000u;
        if (f == 0 || !((live >> w) & 1u)) continue;
        if (j0 + w + L > slen) continue;  // window runs past the sequence (cscore.c:340)
        if (f & 0x80000000u) {
            unsigned long long slot = atomicAdd(P.counters + 0, 1ull);
            if ((int64_t) slot < P.cand_cap) P.cand[slot] = make_key(sorted_idx, p0 + w, 0);
        }
        if (f & 0x00008000u) {
            unsigned long long slot = atomicAdd(P.counters + 0, 1ull);
            if ((int64_t) slot < P.cand_cap) P.cand[slot] = make_key(sorted_idx, p0 + w, 1);
        }
    }
}

template <int G, int W>
__device__ __forceinline__ void run_group(const PrefilterParams &P,
                                          const unsigned char *__restrict__ &tab,
                                          uint32_t &sorted_idx,
                                          const uint32_t (&idx)[PF<W>::kIdx], uint32_t live,
                                          int64_t p0, int32_t j0, int32_t slen) {
    const int cnt = P.batch.g_count[G];
    for (int i = 0; i < cnt; i++) {
        uint32_t acc[W];
        if (score_motif<G, W>(tab, idx, acc))
            emit_candidates<W>(P, acc, live, sorted_idx, p0, j0, slen);
        tab += G * kGroupBytes;
        sorted_idx++;
    }
}

template <int G, int W> struct GroupLoop {
    static __device__ __forceinline__ void run(const PrefilterParams &P,
                                               const unsigned char *__restrict__ &tab,
                                               uint32_t &sorted_idx,
                                               const uint32_t (&idx)[PF<W>::kIdx],
                                               uint32_t live, int64_t p0, int32_t j0,
                                               int32_t slen) {
        GroupLoop<G - 1, W>::run(P, tab, sorted_idx, idx, live, p0, j0, slen);
        run_group<G, W>(P, tab, sorted_idx, idx, live, p0, j0, slen);
    }
};
template <int W> struct GroupLoop<0, W> {
    static __device__ __forceinline__ void run(const PrefilterParams &, const unsigned char *__restrict__ &,
                                               uint32_t &, const uint32_t (&)[PF<W>::kIdx],
                                               uint32_t, int64_t, int32_t, int32_t) {}
};

template <int W>
__global__ void __launch_bounds__(PF<W>::kThreads, 1)
prefilter_kernel(const __grid_constant__ PrefilterParams P) {
    extern __shared__ __align__(16) unsigned char s_tab[];
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(P.tab + P.batch.tab_word_off);
        uint4 *dst = reinterpret_cast<uint4 *>(s_tab);
        const uint32_t n4 = P.batch.tab_words >> 2;
        for (uint32_t i = threadIdx.x; i < n4; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    __syncthreads();

    const SeqView &S = P.seq;
    const int lane = threadIdx.x & 31;
    const int64_t warps_total = (int64_t) gridDim.x * (blockDim.x >> 5);
    const int64_t warp_global = (int64_t) blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t n_chunks = (S.total_packed + PF<W>::kChunk - 1) / PF<W>::kChunk;
    const uint32_t horizon = P.lmax_all >= 32 ? 0xffffffffu : ((1u << P.lmax_all) - 1u);

    for (int64_t chunk = warp_global; chunk < n_chunks; chunk += warps_total) {
        const int64_t p0 = chunk * PF<W>::kChunk + (int64_t) lane * W;
        if (p0 >= S.total_packed) continue;
        const int64_t s = find_seq(S.poff, S.n_seqs, p0);
        const int32_t j0 = (int32_t) (p0 - __ldg(S.poff + s));
        const int32_t slen = __ldg(S.len + s);
        if (j0 >= slen) continue;  // padding behind the sequence's last base

        // N mask of [p0, p0 + W + 31): window w is "dirty" if any base in its horizon is N.
        const uint32_t *mp = S.nmask + (p0 >> 5);
        const uint64_t m64 = (((uint64_t) __ldg(mp + 1) << 32) | __ldg(mp)) >> (p0 & 31);
        uint32_t live = 0;  // windows this thread scores with the table path
#pragma unroll
        for (int w = 0; w < W; w++) {
            const uint32_t bits = (uint32_t) (m64 >> w) & horizon;
            if (j0 + w < slen) {
                if (bits == 0) {
                    live |= 1u << w;
                } else if (P.emit_dirty && !(bits == horizon && !P.any_zero_hit)) {
                    // Touches an N: scored exactly by exact_dirty_kernel.  A window whose whole
                    // horizon is N scores 0 for every motif and is dropped when no motif's
                    // cutoff admits a zero score.
                    unsigned long long slot = atomicAdd(P.counters + 1, 1ull);
                    if ((int64_t) slot < P.dirty_cap) P.dirty[slot] = p0 + w;
                }
            }
        }
        if (live == 0) continue;

        // 2-bit codes of [p0, p0 + 48 - (p0 & 15)) aligned to bit 0.
        const uint32_t *cp = S.codes + (p0 >> 4);
        const uint32_t sh = (uint32_t) (p0 & 15) * 2;
        const uint32_t r0 = __ldg(cp), r1 = __ldg(cp + 1), r2 = __ldg(cp + 2);
        uint32_t cw[4];
        cw[0] = __funnelshift_r(r0, r1, sh);
        cw[1] = __funnelshift_r(r1, r2, sh);
        cw[2] = r2 >> sh;
        cw[3] = 0;
        uint32_t idx[PF<W>::kIdx];
#pragma unroll
        for (int k = 0; k < PF<W>::kIdx; k++) {
            const uint32_t v = __funnelshift_r(cw[(2 * k) >> 5], cw[((2 * k) >> 5) + 1], (2 * k) & 31);
            idx[k] = (v & 0xFu) << 2;
        }

        const unsigned char *tab = s_tab;
        uint32_t sorted_idx = P.batch.first_sorted;
        GroupLoop<kMaxGroups, W>::run(P, tab, sorted_idx, idx, live, p0, j0, slen);
    }
}

// ---------------------------------------------------------------------------------------------
// exact stage: the reference's arithmetic, verbatim (cscore.c:341-389).
// ---------------------------------------------------------------------------------------------
struct ExactParams {
    SeqView seq;
    MotifView mot;
    int strand;
    int key_shift;        // make_site_key's motif shift for this sequence set
    uint64_t *hit_key;    // motif id in the motif field
    double *hit_score;
    int64_t hit_cap;
    unsigned long long *counters;  // [2] hits
};

// Raw score of one strand of window [p, p+L): double accumulation in ascending column order,
// non-ACGT bases skipped.  No multiply on the path, so the sum is the reference's bit for bit.
__device__ __forceinline__ double exact_raw(const SeqView &S, const double *__restrict__ pw, int L,
                                            int64_t p, int rev) {
    double acc = 0.0;
    for (int c = 0; c < L; c++) {
        const int row = base_code(S, p + c);
        if (row >= 0) {
            const double v = rev ? __ldg(pw + 4 * (L - 1 - c) + (3 - row)) : __ldg(pw + 4 * c + row);
            acc = __dadd_rn(acc, v);
        }
    }
    return acc;
}

__device__ __forceinline__ void test_and_emit(const ExactParams &E, uint32_t m, int64_t p, int rev,
                                              double raw) {
    const double score = __ddiv_rn(raw, __ldg(E.mot.max_raw + m));       // cscore.c:357,374
    if (__dsub_rn(score, __ldg(E.mot.cutoff + m)) >= -1e-10) {           // cscore.c:358,375
        unsigned long long slot = atomicAdd(E.counters + 2, 1ull);
        if ((int64_t) slot < E.hit_cap) {
            E.hit_key[slot] = make_site_key(m, p, (uint32_t) rev, E.key_shift);
            E.hit_score[slot] = score;
        }
    }
}

// The bases of window [p, p + 32) as registers: 2-bit codes in a 64-bit word, N flags in 32 bits.
struct Window32 {
    uint64_t codes;
    uint32_t nmask;
};
__device__ __forceinline__ Window32 load_window32(const SeqView &S, int64_t p) {
    // the packed buffers carry >= 4 words of zero padding behind the last sequence (msb_seqs_from_ascii)
    const uint32_t *cp = S.codes + (p >> 4);
    const uint32_t sh = (uint32_t) (p & 15) * 2;
    const uint32_t c0 = __ldg(cp), c1 = __ldg(cp + 1), c2 = __ldg(cp + 2);
    Window32 w;
    w.codes = ((uint64_t) __funnelshift_r(c1, c2, sh) << 32) | __funnelshift_r(c0, c1, sh);
    const uint32_t *mp = S.nmask + (p >> 5);
    w.nmask = __funnelshift_r(__ldg(mp), __ldg(mp + 1), (uint32_t) (p & 31));
    return w;
}
// exact_raw over a register window (L <= 32): same additions in the same order.
__device__ __forceinline__ double exact_raw_w(const Window32 &w, const double *__restrict__ pw, int L, int rev) {
    double acc = 0.0;
    for (int c = 0; c < L; c++) {
        if ((w.nmask >> c) & 1u) continue;
        const int row = (int) ((w.codes >> (2 * c)) & 3u);
        const double v = rev ? __ldg(pw + 4 * (L - 1 - c) + (3 - row)) : __ldg(pw + 4 * c + row);
        acc = __dadd_rn(acc, v);
    }
    return acc;
}

// `col_info` == nullptr: keys carry (sorted motif index, position, strand) and were validated by
// the table prefilter.  Otherwise they are the tensor-core prefilter's raw keys (tile * 256 +
// column, position): decode the column and drop windows that run past their sequence.
__global__ void __launch_bounds__(256)
exact_candidates_kernel(ExactParams E, const uint64_t *__restrict__ cand, int64_t n_cand,
                        const int32_t *__restrict__ order, const uint32_t *__restrict__ col_info) {
    int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_cand) return;
    const uint64_t k = cand[i];
    const int64_t p = key_pos(k);
    uint32_t sorted = key_motif(k);
    int rev = (int) key_rev(k);
    if (col_info) {
        const uint32_t info = __ldg(col_info + sorted);
        if (info == 0xffffffffu) return;   // padding column of a tile
        sorted = info >> 1;
        rev = (int) (info & 1u);
    }
    const uint32_t m = (uint32_t) __ldg(order + sorted);
    const int L = __ldg(E.mot.len + m);
    if (col_info || E.seq.limit) {
        const int64_t s = __ldg(E.seq.blk_seq + (p >> 5));
        const int64_t j = p - __ldg(E.seq.poff + s);
        if (j + L > (int64_t) __ldg(E.seq.len + s)) return;   // cscore.c:340
        if (E.seq.limit && j >= (int64_t) __ldg(E.seq.limit + s)) return;   // the next chunk owns this start
    }
    const double *pw = E.mot.pwm + 4 * (int64_t) __ldg(E.mot.col_off + m);
    test_and_emit(E, m, p, rev, exact_raw_w(load_window32(E.seq, p), pw, L, rev));   // prefilter motifs have L <= 32
}

// The tensor-core prefilter leaves its candidate records in one buffer per epilogue lane
// (prefilter_tc.cuh, "Candidate emission").  lane_totals_kernel: counters[0] += records held (counts
// clamped to the buffer capacity), counters[3] = max(largest unclamped count) -- beyond the capacity
// the host repeats the prefilter with larger buffers.  Both counters are zero before the launch.
__global__ void __launch_bounds__(1024)
lane_totals_kernel(const uint32_t *__restrict__ lane_count, int32_t n_lanes, int64_t cap,
                   unsigned long long *__restrict__ counters) {
    __shared__ unsigned long long s_sum[32];
    __shared__ unsigned int s_max[32];
    unsigned long long sum = 0;
    unsigned int mx = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_lanes; i += gridDim.x * blockDim.x) {
        const unsigned int c = __ldg(lane_count + i);
        mx = max(mx, c);
        sum += min((unsigned long long) c, (unsigned long long) cap);
    }
    for (int d = 16; d; d >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, d);
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    }
    if ((threadIdx.x & 31) == 0) { s_sum[threadIdx.x >> 5] = sum; s_max[threadIdx.x >> 5] = mx; }
    __syncthreads();
    if (threadIdx.x < 32) {
        sum = s_sum[threadIdx.x];
        mx = s_max[threadIdx.x];
        for (int d = 16; d; d >>= 1) {
            sum += __shfl_xor_sync(0xffffffffu, sum, d);
            mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
        }
        if (threadIdx.x == 0) {
            atomicAdd(counters + 0, sum);
            atomicMax(counters + 3, (unsigned long long) mx);
        }
    }
}

// One warp per lane buffer, one thread per record = one window: its bases are loaded once and every
// flagged column of the chunk's 64 is decoded (col_info: tile column -> sorted motif, strand) and
// re-scored.  A warp's 32 records are contiguous (512 B).
__device__ __forceinline__ void exact_record(const ExactParams &E, const uint4 r, const int32_t *__restrict__ order,
                                             const uint32_t *__restrict__ col_info) {
    const int64_t p = (int64_t) (((uint64_t) (r.y & 0x1ffu) << 32) | r.x);
    const uint32_t col0 = r.y >> 9;
    const int64_t s = __ldg(E.seq.blk_seq + (p >> 5));
    const int64_t j = p - __ldg(E.seq.poff + s);
    const int64_t slen = __ldg(E.seq.len + s);
    if (E.seq.limit && j >= (int64_t) __ldg(E.seq.limit + s)) return;   // the next chunk owns this start
    const Window32 w = load_window32(E.seq, p);
    unsigned long long bits = ((unsigned long long) r.w << 32) | r.z;   // one flat loop: less divergence than word by word
    while (bits) {
        const int i = __ffsll((long long) bits) - 1;   // bit 32 h + 8 q + k <=> column 32 h + 4 k + q
        bits &= bits - 1;
        const uint32_t info = __ldg(col_info + col0 + (i & 32) + 4 * (i & 7) + ((i >> 3) & 3));
        if (info == 0xffffffffu) continue;   // padding column of a tile
        const uint32_t m = (uint32_t) __ldg(order + (info >> 1));
        const int rev = (int) (info & 1u);
        const int L = __ldg(E.mot.len + m);
        if (j + L > slen) continue;   // cscore.c:340
        const double *pw = E.mot.pwm + 4 * (int64_t) __ldg(E.mot.col_off + m);
        // the register window holds 32 bases; a longer motif (prefiltered on its first 32 columns) reads memory
        test_and_emit(E, m, p, rev, L <= kMaxFastLen ? exact_raw_w(w, pw, L, rev) : exact_raw(E.seq, pw, L, p, rev));
    }
}

__global__ void __launch_bounds__(256)
exact_records_kernel(ExactParams E, const uint4 *__restrict__ rec, int64_t rec_cap,
                     const uint32_t *__restrict__ lane_count, int32_t n_lanes,
                     const int32_t *__restrict__ order, const uint32_t *__restrict__ col_info) {
    const int64_t b = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (b >= n_lanes) return;
    const int64_t n = min((int64_t) __ldg(lane_count + b), rec_cap);
    const uint4 *buf = rec + b * rec_cap;
    for (int64_t k = threadIdx.x & 31; k < n; k += 32) exact_record(E, __ldg(buf + k), order, col_info);
}

// One thread per (listed position, motif); consecutive threads take consecutive motifs of the
// same position.  `motif_ids` == nullptr means all motifs.
__global__ void __launch_bounds__(256)
exact_positions_kernel(ExactParams E, const int64_t *__restrict__ pos, int64_t n_pos, int64_t pos_base,
                       const int32_t *__restrict__ motif_ids, int32_t n_ids) {
    const int64_t t = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t d = t / n_ids;
    if (d >= n_pos) return;
    const uint32_t m = motif_ids ? (uint32_t) __ldg(motif_ids + (t - d * n_ids)) : (uint32_t) (t - d * n_ids);
    const int64_t p = pos ? __ldg(pos + d) : pos_base + d;
    if (p >= E.seq.total_packed) return;
    const int64_t s = find_seq(E.seq.poff, E.seq.n_seqs, p);
    const int64_t j = p - __ldg(E.seq.poff + s);
    const int L = __ldg(E.mot.len + m);
    const int64_t slen = __ldg(E.seq.len + s);
    if (j >= slen || j + L > slen) return;   // cscore.c:337,340
    if (E.seq.limit && j >= (int64_t) __ldg(E.seq.limit + s)) return;
    const double *pw = E.mot.pwm + 4 * (int64_t) __ldg(E.mot.col_off + m);
    if (E.strand & 1) test_and_emit(E, m, p, 0, exact_raw(E.seq, pw, L, p, 0));
    if (E.strand & 2) test_and_emit(E, m, p, 1, exact_raw(E.seq, pw, L, p, 1));
}

// Dirty windows (they touch a non-ACGT base) x the prefilter-path motifs.  A block takes a chunk of
// kDirtyMotifs motifs, whose fp32 matrices it stages in shared memory, and 8 x 32 listed positions:
// one position per lane (the window's bases in registers), each warp walking the chunk's motifs
// together, so the motif, its length and the column are warp-uniform and the 32 lanes' reads of one
// column hit at most 4 words of one 16-byte line.  (Reading the matrices straight from global memory
// left one sector in flight per warp: 0.71 ms for 26.6 k positions x 750 motifs; staged: see
// profiles/.)  Windows are screened in fp32 (upload_floors in msb200.cu); the few that can reach the
// cutoff are re-scored with the reference's fp64 arithmetic from global memory.
constexpr int kDirtyMotifs = 64;
constexpr int kDirtyStride = 4 * kMaxFastLen;   // floats per staged motif
__global__ void __launch_bounds__(256)
exact_dirty_kernel(ExactParams E, const int64_t *__restrict__ pos, int64_t n_pos,
                   const int32_t *__restrict__ motif_ids, int32_t n_ids) {
    __shared__ float s_pw[kDirtyMotifs * kDirtyStride];
    __shared__ float s_floor[kDirtyMotifs];
    __shared__ int32_t s_len[kDirtyMotifs];
    __shared__ uint32_t s_motif[kDirtyMotifs];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int32_t n_chunks = (n_ids + kDirtyMotifs - 1) / kDirtyMotifs;
    const int32_t chunk = (int32_t) (blockIdx.x % (unsigned) n_chunks);
    const int64_t group = (int64_t) (blockIdx.x / (unsigned) n_chunks) * 8 + wib;
    const int32_t k0 = chunk * kDirtyMotifs, n_here = min(n_ids - k0, kDirtyMotifs);
    for (int32_t kl = wib; kl < n_here; kl += 8) {
        const uint32_t m = (uint32_t) __ldg(motif_ids + k0 + kl);
        const int L = __ldg(E.mot.len + m);
        const float *src = E.mot.pwm32 + 4 * (int64_t) __ldg(E.mot.col_off + m);
        if (L <= kMaxFastLen)   // longer motifs are not screened: they go straight to the fp64 path below
            for (int i = lane; i < 4 * L; i += 32) s_pw[kl * kDirtyStride + i] = __ldg(src + i);
        if (lane == 0) { s_len[kl] = L; s_floor[kl] = __ldg(E.mot.floor32 + m); s_motif[kl] = m; }
    }
    __syncthreads();
    if (group * 32 >= n_pos) return;
    const int64_t d = group * 32 + lane;
    int64_t p = 0, left = 0;
    Window32 w;
    w.codes = 0;
    w.nmask = 0;
    if (d < n_pos) {
        p = __ldg(pos + d);
        const int64_t s = __ldg(E.seq.blk_seq + (p >> 5));
        const int64_t j = p - __ldg(E.seq.poff + s);
        left = (int64_t) __ldg(E.seq.len + s) - j;
        if (E.seq.limit && j >= (int64_t) __ldg(E.seq.limit + s)) left = 0;   // the next chunk owns this start
        if (left > 0) w = load_window32(E.seq, p);
    }
    for (int32_t kl = 0; kl < n_here; kl++) {
        const int L = s_len[kl];
        if (L > kMaxFastLen) {   // warp-uniform: a long motif, scored from memory with the reference's arithmetic
            if (L <= left) {
                const uint32_t m = s_motif[kl];
                const double *pw = E.mot.pwm + 4 * (int64_t) __ldg(E.mot.col_off + m);
                if (E.strand & 1) test_and_emit(E, m, p, 0, exact_raw(E.seq, pw, L, p, 0));
                if (E.strand & 2) test_and_emit(E, m, p, 1, exact_raw(E.seq, pw, L, p, 1));
            }
            continue;
        }
        const float *pw32 = s_pw + kl * kDirtyStride;
        // fp32 screen of both strands (FP64 issues at a fraction of the FP32 rate and the division in
        // test_and_emit is a subroutine): a score below the floor cannot pass
        // Unrolled over the columns with compile-time shifts and offsets: per column and lane one bit
        // field extract, two address adds, two shared loads and two predicated adds (the kernel is
        // issue-bound: 7e8 column steps per scan of configs[1]).
        float f32 = 0.f, r32 = 0.f;
        const float *rev_end = pw32 + 4 * (L - 1) + 3;   // reverse strand reads matrix column L - 1 - c, row 3 - row
        const uint32_t lo = (uint32_t) w.codes, hi = (uint32_t) (w.codes >> 32);
#pragma unroll
        for (int c = 0; c < kMaxFastLen; c++) {
            if (c >= L) break;   // warp-uniform
            const uint32_t row = ((c < 16 ? lo : hi) >> (2 * (c & 15))) & 3u;
            const float vf = pw32[4 * c + row], vr = *(rev_end - 4 * c - row);
            if (!((w.nmask >> c) & 1u)) { f32 += vf; r32 += vr; }
        }
        if (L > left) continue;   // cscore.c:337,340 (also lanes without a position: left = 0)
        const float floor32 = s_floor[kl];
        const bool try_f = (E.strand & 1) && !(f32 < floor32), try_r = (E.strand & 2) && !(r32 < floor32);
        if (try_f || try_r) {   // rare: the reference's arithmetic, matrices from global memory
            const uint32_t m = s_motif[kl];
            const double *pw = E.mot.pwm + 4 * (int64_t) __ldg(E.mot.col_off + m);
            if (try_f) test_and_emit(E, m, p, 0, exact_raw_w(w, pw, L, 0));
            if (try_r) test_and_emit(E, m, p, 1, exact_raw_w(w, pw, L, 1));
        }
    }
}

__global__ void __launch_bounds__(256)
decode_sites_kernel(SeqView S, const uint64_t *__restrict__ key, int64_t n, int key_shift,
                    int32_t *__restrict__ seq_idx, int32_t *__restrict__ start,
                    int8_t *__restrict__ strand) {
    int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t k = key[i];
    const int64_t p = site_pos(k, key_shift);
    const int64_t s = __ldg(S.blk_seq + (p >> 5));   // a site's position lies inside a sequence, so its block has an owner
    seq_idx[i] = (int32_t) s;
    start[i] = (int32_t) (p - __ldg(S.poff + s));
    strand[i] = (int8_t) (key_rev(k) + 1);  // 1 forward, 2 reverse (cscore.c:359,376)
}

// MSB_SCAN_COMPACT: a site's packed position and strand are the low key_shift (<= 32) bits of its sorted key
// (make_site_key); 4 bytes per site cross PCIe instead of 9 and the host finds the owning sequence.
__global__ void __launch_bounds__(256)
compact_sites_kernel(const uint64_t *__restrict__ key, int64_t n, int key_shift, uint32_t *__restrict__ out) {
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint32_t) (key[i] & ((1ull << key_shift) - 1ull));
}

// offsets[m] = first sorted site whose motif id is >= m  (m = 0..n_motifs): CSR over motifs.
__global__ void __launch_bounds__(256)
motif_offsets_kernel(const uint64_t *__restrict__ key, int64_t n, int32_t n_motifs, int key_shift,
                     int64_t *__restrict__ offsets) {
    const int32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m > n_motifs) return;
    int64_t lo = 0, hi = n;  // first i with site_motif(key[i]) >= m
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if ((int64_t) site_motif(__ldg(key + mid), key_shift) < (int64_t) m) lo = mid + 1; else hi = mid;
    }
    offsets[m] = lo;
}

// MSB_SCAN_COUNTS: per-motif site counts straight from the unsorted hit keys (a block-private histogram in
// shared memory when the motif set fits, flushed with one global atomic per non-empty bin).
__global__ void __launch_bounds__(256)
count_hits_kernel(const uint64_t *__restrict__ key, int64_t n, int key_shift, int32_t n_motifs, int in_smem,
                  unsigned long long *__restrict__ counts) {
    extern __shared__ unsigned int s_hist[];
    if (in_smem) {
        for (int32_t m = threadIdx.x; m < n_motifs; m += blockDim.x) s_hist[m] = 0;
        __syncthreads();
    }
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t) gridDim.x * blockDim.x) {
        const uint32_t m = site_motif(__ldg(key + i), key_shift);
        if (in_smem) atomicAdd(s_hist + m, 1u); else atomicAdd(counts + m, 1ull);
    }
    if (in_smem) {
        __syncthreads();
        for (int32_t m = threadIdx.x; m < n_motifs; m += blockDim.x)
            if (s_hist[m]) atomicAdd(counts + m, (unsigned long long) s_hist[m]);
    }
}

// Per motif, the number of sequences with at least one site (what motif_enrichment counts,
// stats.py:29-31), from the sorted site list: a site opens a new (motif, sequence) cell if it is the
// first of its motif or its sequence differs from the previous site's.
__global__ void __launch_bounds__(256)
region_counts_kernel(const uint64_t *__restrict__ key, const int32_t *__restrict__ seq_idx, int64_t n,
                     int key_shift, unsigned long long *__restrict__ out) {
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t m = site_motif(key[i], key_shift);
    if (i == 0 || site_motif(key[i - 1], key_shift) != m || seq_idx[i - 1] != seq_idx[i]) atomicAdd(out + m, 1ull);
}

// ---------------------------------------------------------------------------------------------
// De-duplication of adjacent sites (scanner.py:156-193), on the sorted site list.
// Per (motif, sequence) segment and per strand separately, walking sites by ascending start with
// one "current survivor": a site closer than the motif length to the survivor replaces it only
// if its score is strictly higher (`curr.score >= next.score` keeps curr, scanner.py:163).
// The surviving forward and reverse sites keep their (start, forward-first) order, which is the
// reference's `fwd + rev` followed by a stable sort on start (scanner.py:190-191).
//
// The walk only carries state across same-strand sites that are closer than the motif length to their
// predecessor (the survivor is never behind the predecessor's predecessor chain), so a segment falls
// apart into independent CHAINS at every same-strand gap >= L.  One thread per site: it looks back (at most
// the few sites within L positions) to see whether it opens a chain, and if so walks that chain.  Whole
// chromosomes de-duplicate as fast as 1 kb regions; only a run of overlapping sites is sequential.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dedup_flags_kernel(const uint64_t *__restrict__ key, const int32_t *__restrict__ seq_idx,
                   const int32_t *__restrict__ start, const int8_t *__restrict__ strand,
                   const double *__restrict__ score, int64_t n, const int32_t *__restrict__ mlen, int key_shift,
                   int32_t *__restrict__ keep) {
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t m = site_motif(key[i], key_shift);
    const int32_t s = seq_idx[i];
    const int32_t L = __ldg(mlen + m);
    const int8_t which = strand[i];
    for (int64_t t = i - 1; t >= 0; t--) {   // does a same-strand site of this segment sit less than L before this one?
        if (site_motif(key[t], key_shift) != m || seq_idx[t] != s) break;
        if (start[i] - start[t] >= L) break;
        if (strand[t] == which) return;      // not a chain head: the head's thread reaches this site
    }
    int64_t cur = i, last = i;
    keep[i] = 1;
    for (int64_t t = i + 1; t < n; t++) {
        if (site_motif(key[t], key_shift) != m || seq_idx[t] != s) break;
        if (strand[t] != which) continue;
        if (start[t] - start[last] >= L) break;   // the next chain
        last = t;
        if (start[t] - start[cur] < L) {
            if (score[cur] >= score[t]) { keep[t] = 0; continue; }
            keep[cur] = 0;
        }
        keep[t] = 1;
        cur = t;
    }
}

__global__ void __launch_bounds__(256)
scatter_kept_kernel(const int32_t *__restrict__ keep, const int64_t *__restrict__ dst, int64_t n,
                    const uint64_t *__restrict__ key, const double *__restrict__ score,
                    const int32_t *__restrict__ seq_idx, const int32_t *__restrict__ start,
                    const int8_t *__restrict__ strand, uint64_t *__restrict__ o_key,
                    double *__restrict__ o_score, int32_t *__restrict__ o_seq,
                    int32_t *__restrict__ o_start, int8_t *__restrict__ o_strand) {
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !keep[i]) return;
    const int64_t j = dst[i];
    o_key[j] = key[i];
    o_score[j] = score[i];
    o_seq[j] = seq_idx[i];
    o_start[j] = start[i];
    o_strand[j] = strand[i];
}

// ---------------------------------------------------------------------------------------------
// c_score: offset-0 window of every sequence (cscore.c:191-223).  One thread per sequence,
// looping over a slice of motifs; out[m * n_seqs + i].
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
score0_kernel(SeqView S, MotifView M, int strand, int32_t m_begin, int32_t m_end, int64_t out_stride,
              double *__restrict__ out) {
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S.n_seqs) return;
    const int64_t p = __ldg(S.poff + i);
    for (int32_t m = m_begin + (int32_t) blockIdx.y; m < m_end; m += (int32_t) gridDim.y) {
        const int L = __ldg(M.len + m);
        const double *pw = M.pwm + 4 * (int64_t) __ldg(M.col_off + m);
        double f = 0.0, r = 0.0;
        for (int c = 0; c < L; c++) {
            const int row = base_code(S, p + c);
            if (row >= 0) {
                if (strand & 1) f = __dadd_rn(f, __ldg(pw + 4 * c + row));
                if (strand & 2) r = __dadd_rn(r, __ldg(pw + 4 * (L - 1 - c) + (3 - row)));
            }
        }
        double sc = 0.0;
        if (strand == 1) sc = f;
        else if (strand == 2) sc = r;
        else if (strand == 3) sc = (f > r) ? f : r;  // cscore.c:215-221
        out[(int64_t) (m - m_begin) * out_stride + i] = __ddiv_rn(sc, __ldg(M.max_raw + m));
    }
}

// ---------------------------------------------------------------------------------------------
// Order statistics of score distributions (get_score_cutoffs, motif/__init__.py:378-401: sort
// descending, read a handful of indices) by radix SELECT instead of a sort.
//
// order_key: the bijection double -> uint64 whose unsigned order is the order a radix sort of doubles
// uses (negative values below positive ones, -0.0 below +0.0, NaNs at the ends by bit pattern), so the
// selected values are the ones msb_score_select's former cub::DeviceSegmentedRadixSort returned.
// Key 0 (the bit pattern of a negative NaN with every payload bit set) marks a dead entry.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t order_key(double v) {
    const uint64_t b = (uint64_t) __double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double order_key_inv(uint64_t k) {
    const uint64_t b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long) b);
}

constexpr int kSelMaxRanks = 8;
constexpr int kSelThreads = 512;

// One block per segment.  Segment s holds the keys (kFromDoubles: doubles, keyed on the fly) at
// data[begin .. end) with begin = seg_off ? seg_off[s] : s * stride and end = seg_off ? seg_off[s + 1] :
// begin + stride.  out[s * n_ranks + j] = the ranks[j]-th largest value (0-based), found digit by digit
// from the top: per 8-bit digit one pass over the segment builds, for every group of ranks that still share
// a prefix, the histogram of the next digit among the keys with that prefix; the bucket that holds the rank
// extends the prefix.  Eight passes for all ranks together; the segment is staged in shared memory when
// it fits (smem_keys), else re-read from L2.  Warp-level vote merges equal digits before the
// shared-memory atomic: score keys share their top bytes.  ok[s] (optional) = 0 when some rank is not
// below live[s] (the number of entries that are not dead).
template <bool kFromDoubles>
__global__ void __launch_bounds__(kSelThreads)
select_ranks_kernel(const void *__restrict__ data, const int64_t *__restrict__ seg_off, int64_t stride,
                    const int64_t *__restrict__ ranks, int n_ranks, int smem_keys, const int64_t *__restrict__ live,
                    double *__restrict__ out, int32_t *__restrict__ ok) {
    extern __shared__ __align__(16) uint64_t s_keys[];
    __shared__ unsigned int s_hist[kSelMaxRanks][256];
    __shared__ uint64_t s_prefix[kSelMaxRanks];
    __shared__ long long s_rank[kSelMaxRanks];
    __shared__ int s_group[kSelMaxRanks], s_ngroups;
    const int seg = blockIdx.x;
    const int64_t begin = seg_off ? seg_off[seg] : (int64_t) seg * stride;
    const int64_t n = seg_off ? seg_off[seg + 1] - begin : stride;
    const int64_t n_live = live ? live[seg] : n;
    const uint64_t *keys = reinterpret_cast<const uint64_t *>(data) + begin;
    const double *vals = reinterpret_cast<const double *>(data) + begin;
    const bool staged = n <= (int64_t) smem_keys;
    if (staged)
        for (int64_t i = threadIdx.x; i < n; i += kSelThreads) s_keys[i] = kFromDoubles ? order_key(vals[i]) : keys[i];
    if (threadIdx.x < n_ranks) {
        s_prefix[threadIdx.x] = 0;
        s_rank[threadIdx.x] = ranks[threadIdx.x];
    }
    if (threadIdx.x == 0) {
        bool fine = true;
        for (int j = 0; j < n_ranks; j++) fine = fine && ranks[j] >= 0 && ranks[j] < n_live;
        if (ok) ok[seg] = fine ? 1 : 0;
        s_ngroups = fine ? 1 : 0;
        for (int j = 0; j < n_ranks; j++) s_group[j] = 0;
    }
    __syncthreads();
    if (s_ngroups == 0) {   // not enough live entries: the caller takes another path for this segment
        if (threadIdx.x < n_ranks) out[(int64_t) seg * n_ranks + threadIdx.x] = __longlong_as_double(0x7ff8000000000000ll);
        return;
    }
    for (int shift = 56; shift >= 0; shift -= 8) {
        const int ng = s_ngroups;
        for (int i = threadIdx.x; i < ng * 256; i += kSelThreads) (&s_hist[0][0])[i] = 0;
        __syncthreads();
        const uint64_t mask = shift == 56 ? 0ull : (~0ull << (shift + 8));
        // group g's prefix = the prefix of its first rank
        uint64_t gp[kSelMaxRanks];
        {
            int g = 0;
            for (int j = 0; j < n_ranks && g < ng; j++)
                if (s_group[j] == g) gp[g++] = s_prefix[j];
        }
        const int64_t n_round = (n + 31) / 32 * 32;   // whole warps take part in the vote
        for (int64_t i = threadIdx.x; i < n_round; i += kSelThreads) {
            const bool have = i < n;
            const uint64_t k = !have ? 0ull : (staged ? s_keys[i] : (kFromDoubles ? order_key(vals[i]) : keys[i]));
            int g = -1;
            if (have)
                for (int t = 0; t < ng; t++)
                    if ((k & mask) == gp[t]) g = t;
            // lanes with the same (group, digit) elect one to add their number
            const unsigned int tag = g < 0 ? 0xffffffffu : (unsigned int) (g * 256 + (int) ((k >> shift) & 255));
            const unsigned int peers = __match_any_sync(0xffffffffu, tag);
            if (g >= 0 && (threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&s_hist[0][0] + tag, (unsigned int) __popc(peers));
        }
        __syncthreads();
        if (threadIdx.x < n_ranks) {
            const int j = threadIdx.x;
            const unsigned int *h = s_hist[s_group[j]];
            long long r = s_rank[j], above = 0;
            int b = 255;
            for (; b > 0; b--) {
                if (above + (long long) h[b] > r) break;
                above += h[b];
            }
            s_rank[j] = r - above;
            s_prefix[j] |= (uint64_t) b << shift;
        }
        __syncthreads();
        if (threadIdx.x == 0) {   // regroup: ranks with equal prefixes share a histogram in the next pass
            int ngn = 0;
            for (int j = 0; j < n_ranks; j++) {
                int g = -1;
                for (int t = 0; t < j; t++)
                    if (s_prefix[t] == s_prefix[j]) { g = s_group[t]; break; }
                s_group[j] = g >= 0 ? g : ngn++;
            }
            s_ngroups = ngn;
        }
        __syncthreads();
    }
    if (threadIdx.x < n_ranks) out[(int64_t) seg * n_ranks + threadIdx.x] = order_key_inv(s_prefix[threadIdx.x]);
}

// After a scan with a start limit of 1 (only the offset-0 window of every sample) the sorted site list holds,
// per motif, one site per (sample, strand) that reached the pilot cutoff.  c_score's value for strand 3 is the
// better strand (cscore.c:215-221): key[i] = order_key(max over the sample's sites); the second site of a
// sample becomes a dead entry (key 0) and is counted in dead[motif].
__global__ void __launch_bounds__(256)
sample_keys_kernel(const uint64_t *__restrict__ site_key, const int32_t *__restrict__ seq_idx,
                   const double *__restrict__ score, int64_t n, int key_shift, uint64_t *__restrict__ key,
                   unsigned long long *__restrict__ dead) {
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t m = site_motif(site_key[i], key_shift);
    const int32_t s = seq_idx[i];
    if (i > 0 && site_motif(site_key[i - 1], key_shift) == m && seq_idx[i - 1] == s) {
        key[i] = 0;
        atomicAdd(dead + m, 1ull);
        return;
    }
    uint64_t k = order_key(score[i]);
    if (i + 1 < n && site_motif(site_key[i + 1], key_shift) == m && seq_idx[i + 1] == s) k = max(k, order_key(score[i + 1]));
    key[i] = k;
}

// live[m] = offsets[m + 1] - offsets[m] - dead[m]
__global__ void __launch_bounds__(256)
live_counts_kernel(const int64_t *__restrict__ offsets, const unsigned long long *__restrict__ dead, int32_t n_motifs,
                   int64_t *__restrict__ live) {
    const int32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m < n_motifs) live[m] = offsets[m + 1] - offsets[m] - (int64_t) dead[m];
}

// fill an int32 array with one value (the start limit of the offset-0 scan)
__global__ void __launch_bounds__(256) fill_i32_kernel(int32_t *__restrict__ p, int64_t n, int32_t v) {
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

}  // namespace msb
