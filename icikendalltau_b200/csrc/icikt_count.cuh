// icikt_count.cuh -- the memory-space accessors and pass A, the bucket-free inversion-counting
// pass on 16-bit keys (see icikt_pairs.cu for the counting scheme).  Shared by the pair kernel
// (icikt_pairs.cu) and the fused column kernel (icikt_columns.cu), which runs the pass on the
// column's own sorted ranks to get the per-column constant `cconst`.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

// unroll factors of pass A's two sweeps over a thread's 8-key runs (tuning knobs; 1 = one run per iteration)
#ifndef ICIKT_COUNT_UNROLL
#define ICIKT_COUNT_UNROLL 1
#endif
#ifndef ICIKT_HI_KEY_MASKS
#define ICIKT_HI_KEY_MASKS 1
#endif
#ifndef ICIKT_SCATTER_UNROLL
#define ICIKT_SCATTER_UNROLL 1
#endif

namespace icikt {
namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int kCountUnroll = ICIKT_COUNT_UNROLL, kScatterUnroll = ICIKT_SCATTER_UNROLL;

__device__ __forceinline__ uint32_t lanemask_lt() {
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}
__device__ __forceinline__ uint32_t lanemask_le() {
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_le;" : "=r"(m));
  return m;
}
// lanes strictly below the highest set bit of `bits` (bits != 0)
__device__ __forceinline__ uint32_t below_top(uint32_t bits) {
  return (1u << (31 - __clz((int)bits))) - 1u;
}

// ---- memory-space accessors ----------------------------------------------------------------
// The counting passes run either on shared memory (32-bit shared-window addresses, explicit
// ld/st.shared) or, for vectors too long for one CTA's shared memory, on a per-CTA scratch in
// global memory that stays L2-resident.  Offsets are in bytes.
__device__ __forceinline__ uint32_t smem_addr(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
// Constants handed over as kernel parameters so that the assembler keeps multiplications by
// them as IMAD / IMAD.HI (FMA pipe) instead of strength-reducing them to integer-ALU shifts/adds.
struct PipeConst {
  uint32_t one, two, c64k;
  uint32_t c4410, c4432;  // PRMT selectors: low / high 16-bit half of a register, zero-extended
};
__host__ __device__ inline PipeConst make_pipe_const() {
  PipeConst pc;
  pc.one = 1u;
  pc.two = 2u;
  pc.c64k = 65536u;
  pc.c4410 = 0x4410u;
  pc.c4432 = 0x4432u;
  return pc;
}

// ---- TMA bulk copy global -> shared, completion on an mbarrier (sm_90+) ---------------------
// Used for the one contiguous bulk transfer of the pair kernel: y's dense-rank table into the
// second sequence buffer.  One thread issues it; the copy engine moves the bytes while the CTA
// does the bit-mask counting, then everybody waits on the barrier's phase.
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
  // earlier generic-proxy accesses to the destination are ordered before the async-proxy writes
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(mbar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(mbar), "r"(parity)
        : "memory");
  } while (!done);
}

template <bool G>
struct Mem;
template <>
struct Mem<false> {
  typedef uint32_t ptr;
  static __device__ __forceinline__ ptr add(ptr p, int32_t bytes) { return p + (uint32_t)bytes; }
  static __device__ __forceinline__ uint32_t ld16(ptr p) {
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(p) : "memory");
    return v;
  }
  // 16-bit load zero-extended straight into a 32-bit register (no conversion afterwards)
  static __device__ __forceinline__ uint32_t ld16w(ptr p) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(p) : "memory");
    return v;
  }
  static __device__ __forceinline__ uint32_t ld32(ptr p) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(p) : "memory");
    return v;
  }
  static __device__ __forceinline__ void ld128(ptr p, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(p) : "memory");
  }
  // position of p for the inversion accumulator: any value that differs from the byte offset
  // inside the buffer by a per-level constant (here the shared-window address itself)
  static __device__ __forceinline__ uint32_t off(ptr p, ptr) { return p; }
  // a + b on the FMA pipe (IMAD with a multiplier the assembler cannot fold): the integer ALU
  // pipe is the binding one in pass A, both pipes issue one warp instruction per two cycles
  static __device__ __forceinline__ uint32_t fadd(uint32_t a, uint32_t b, uint32_t one) {
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(one), "r"(b));
    return d;
  }
  // high 16-bit key of a word on the FMA pipe (IMAD.HI by 65536)
  static __device__ __forceinline__ uint32_t hi16(uint32_t w, uint32_t c64k) {
    uint32_t d;
    asm("mul.hi.u32 %0, %1, %2;" : "=r"(d) : "r"(w), "r"(c64k));
    return d;
  }
  // one key of pass A's scatter sweep (two bits per level, see count_pass): Q01/Q23 hold the
  // next free slot (element index, 16 bits each) of digit classes 0|1 and 2|3.  acc2 sums, over
  // ALL keys, the running slot of the key's lo = 1 sibling class (one IMAD.HI on the FMA pipe, no
  // predicate); count_pass takes the lo = 1 keys' own share back out in closed form.  `two` and the
  // other constants are registers the assembler cannot fold, so the slot address is an IMAD and
  // not an LEA.  11 instructions per key.
  static __device__ __forceinline__ void step4(uint32_t& Q01, uint32_t& Q23, uint32_t& acc2, uint32_t w,
                                               uint32_t bit_lo, uint32_t bit_hi, uint32_t key, ptr base,
                                               const PipeConst& pc) {
    asm volatile(
        "{\n\t"
        ".reg .pred ph, pl;\n\t"
        ".reg .b32 t, qs, sel, idx, ad, inc;\n\t"
        "and.b32 t, %3, %5;\n\t"
        "setp.ne.u32 ph, t, 0;\n\t"
        "and.b32 t, %3, %4;\n\t"
        "setp.ne.u32 pl, t, 0;\n\t"
        "selp.b32 qs, %1, %0, ph;\n\t"
        "mad.hi.u32 %2, qs, %9, %2;\n\t"
        "selp.b32 sel, %11, %10, pl;\n\t"
        "prmt.b32 idx, qs, 0, sel;\n\t"
        "mad.lo.u32 ad, idx, %8, %7;\n\t"
        "st.shared.u16 [ad], %6;\n\t"
        "selp.b32 inc, %9, %12, pl;\n\t"
        "@ph add.u32 %1, %1, inc;\n\t"
        "@!ph add.u32 %0, %0, inc;\n\t"
        "}"
        : "+r"(Q01), "+r"(Q23), "+r"(acc2)
        : "r"(w), "r"(bit_lo), "r"(bit_hi), "h"((unsigned short)key), "r"(base), "r"(pc.two), "r"(pc.c64k),
          "r"(pc.c4410), "r"(pc.c4432), "r"(pc.one)
        : "memory");
  }
  // the same without the store: the last level only has to count
  static __device__ __forceinline__ void step4c(uint32_t& Q01, uint32_t& Q23, uint32_t& acc2, uint32_t w,
                                                uint32_t bit_lo, uint32_t bit_hi, const PipeConst& pc) {
    asm volatile(
        "{\n\t"
        ".reg .pred ph, pl;\n\t"
        ".reg .b32 t, qs, inc;\n\t"
        "and.b32 t, %3, %5;\n\t"
        "setp.ne.u32 ph, t, 0;\n\t"
        "and.b32 t, %3, %4;\n\t"
        "setp.ne.u32 pl, t, 0;\n\t"
        "selp.b32 qs, %1, %0, ph;\n\t"
        "mad.hi.u32 %2, qs, %6, %2;\n\t"
        "selp.b32 inc, %6, %7, pl;\n\t"
        "@ph add.u32 %1, %1, inc;\n\t"
        "@!ph add.u32 %0, %0, inc;\n\t"
        "}"
        : "+r"(Q01), "+r"(Q23), "+r"(acc2)
        : "r"(w), "r"(bit_lo), "r"(bit_hi), "r"(pc.c64k), "r"(pc.one));
  }
  static __device__ __forceinline__ void red_add32(ptr p, uint32_t v) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(p), "r"(v) : "memory");
  }
  static __device__ __forceinline__ uint32_t atom_or32(ptr p, uint32_t v) {  // returns the old word
    uint32_t old;
    asm volatile("atom.shared.or.b32 %0, [%1], %2;" : "=r"(old) : "r"(p), "r"(v) : "memory");
    return old;
  }
  static __device__ __forceinline__ ptr from_shared(const void* q) { return smem_addr(q); }
  static __device__ __forceinline__ void st16(ptr p, uint32_t v) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(p), "h"((unsigned short)v) : "memory");
  }
  static __device__ __forceinline__ void st32(ptr p, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(p), "r"(v) : "memory");
  }
  static __device__ __forceinline__ void st128(ptr p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
  }
};
template <>
struct Mem<true> {
  typedef unsigned char* ptr;
  static __device__ __forceinline__ ptr add(ptr p, int32_t bytes) { return p + bytes; }
  static __device__ __forceinline__ uint32_t ld16(ptr p) { return *reinterpret_cast<const unsigned short*>(p); }
  static __device__ __forceinline__ uint32_t ld16w(ptr p) { return *reinterpret_cast<const unsigned short*>(p); }
  static __device__ __forceinline__ uint32_t ld32(ptr p) { return *reinterpret_cast<const uint32_t*>(p); }
  static __device__ __forceinline__ void ld128(ptr p, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    a = v.x; b = v.y; c = v.z; d = v.w;
  }
  static __device__ __forceinline__ uint32_t off(ptr p, ptr base) { return (uint32_t)(p - base); }
  static __device__ __forceinline__ uint32_t fadd(uint32_t a, uint32_t b, uint32_t) { return a + b; }
  static __device__ __forceinline__ uint32_t hi16(uint32_t w, uint32_t) { return w >> 16; }
  static __device__ __forceinline__ void step4(uint32_t& Q01, uint32_t& Q23, uint32_t& acc2, uint32_t w,
                                               uint32_t bit_lo, uint32_t bit_hi, uint32_t key, ptr base,
                                               const PipeConst&) {
    const bool ph = (w & bit_hi) != 0u, pl = (w & bit_lo) != 0u;
    const uint32_t qs = ph ? Q23 : Q01;
    const uint32_t idx = pl ? (qs >> 16) : (qs & 0xffffu);
    *reinterpret_cast<unsigned short*>(base + 2 * (size_t)idx) = (unsigned short)key;
    const uint32_t inc = pl ? 0x10000u : 1u;
    if (ph) Q23 += inc; else Q01 += inc;
    acc2 += qs >> 16;
  }
  static __device__ __forceinline__ void step4c(uint32_t& Q01, uint32_t& Q23, uint32_t& acc2, uint32_t w,
                                                uint32_t bit_lo, uint32_t bit_hi, const PipeConst&) {
    const bool ph = (w & bit_hi) != 0u, pl = (w & bit_lo) != 0u;
    const uint32_t qs = ph ? Q23 : Q01;
    const uint32_t inc = pl ? 0x10000u : 1u;
    if (ph) Q23 += inc; else Q01 += inc;
    acc2 += qs >> 16;
  }
  static __device__ __forceinline__ void red_add32(ptr p, uint32_t v) { atomicAdd(reinterpret_cast<uint32_t*>(p), v); }
  static __device__ __forceinline__ uint32_t atom_or32(ptr p, uint32_t v) { return atomicOr(reinterpret_cast<uint32_t*>(p), v); }
  // generic addressing reaches shared memory too
  static __device__ __forceinline__ ptr from_shared(const void* q) { return (ptr) const_cast<void*>(q); }
  static __device__ __forceinline__ void st16(ptr p, uint32_t v) {
    *reinterpret_cast<unsigned short*>(p) = (unsigned short)v;
  }
  static __device__ __forceinline__ void st32(ptr p, uint32_t v) { *reinterpret_cast<uint32_t*>(p) = v; }
  static __device__ __forceinline__ void st128(ptr p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    *reinterpret_cast<uint4*>(p) = make_uint4(a, b, c, d);
  }
};

// Pass A: the bucket-free counting pass on 16-bit keys (see the file header), lane-sequential,
// TWO bits per level.  Thread t of the CTA owns the contiguous range [t*R, (t+1)*R) of the
// sequence, R = 8*kk keys, and walks it with 128-bit loads (kk must be ODD: the 16-byte accesses
// of a quarter warp then fall into distinct bank groups).  A level on bits (s+1, s) is the fusion
// of the two one-bit levels s+1 and s: a stable partition of the whole sequence into the digit
// classes 0,1,2,3 (digit = 2*hi + lo), and the count
//     sum over keys with hi = 0 of #(hi = 1 keys before it)                        [bit s+1]
//   + sum over keys with lo = 0 of #(lo = 1 keys before it after the hi partition)  [bit s]
// where the second term is, for a key of class 0, the class-1 keys before it, and for a key of
// class 2 all N1 class-1 keys plus the class-3 keys before it.  Per level:
//   sweep 1  counts the range's keys per class, two keys per 32-bit word at once
//            ((w >> s) & 0x00010001 summed in two 16-bit fields), and -- same fields, one more
//            IMAD per word -- the sum J1 of the local indices of the hi = 1 keys: the whole hi-bit
//            term of the thread follows from it in closed form (a hi = 0 key at local index j has
//            j - (hi = 0 keys before it) hi = 1 keys before it inside the range);
//   two packed warp scans + warp-total reductions give the class counts before the range;
//   sweep 2  walks the range with four running slot indices (packed 2 x 16 bit in Q01, Q23),
//            stores every key to its slot of the other buffer and sums the running slot of the
//            key's lo = 1 sibling class over all keys (step4); the lo = 1 keys' own share of that
//            sum -- their destination slots, two arithmetic series -- is taken out in closed form.
// No ballots, no per-key population counts: about 16 instructions per key and level of two bits.
// With an odd number of bits the first level treats the missing top bit as zero.
//   acc64 += the count over all levels (this thread's share)
struct LevelCounts {
  uint32_t n1, n2, n3, J1;
};
// class counts and the hi-index sum of one thread's range from the packed sums of sweep 1:
// cl/ch/cb = per-half sums of the lo / hi / lo&hi bits, wm = per-half sums of (word index) * hi bit
__device__ __forceinline__ LevelCounts fold_counts(uint32_t cl, uint32_t ch, uint32_t cb, uint32_t wm) {
  LevelCounts c;
  c.n3 = (cb & 0xffffu) + (cb >> 16);
  c.n2 = (ch & 0xffffu) + (ch >> 16) - c.n3;
  c.n1 = (cl & 0xffffu) + (cl >> 16) - c.n3;
  // key 2m sits in the low half of word m, key 2m+1 in the high half
  c.J1 = 2u * ((wm & 0xffffu) + (wm >> 16)) + (ch >> 16);
  return c;
}
// the thread's share of a level's count (all arithmetic modulo 2^32; the true value is small):
//   hi term   n_hi0 * (hi keys before the range) + sum_{hi=0 keys} j - C(n_hi0, 2)
//   lo term   acc2 (all keys) - slots of the class-1 keys - slots of the class-3 keys
//             - n0 * N0 - n2 * (N0 + N2)     (class bases; N1 class-1 keys precede every class-2 key)
__device__ __forceinline__ uint32_t level_share(const LevelCounts& c, uint32_t R, uint32_t acc2, uint32_t e23,
                                                uint32_t start1, uint32_t start3, uint32_t N0, uint32_t N2) {
  const uint32_t nh0 = R - c.n2 - c.n3, n0 = nh0 - c.n1;
  const uint32_t hi_term = nh0 * e23 + ((R * (R - 1u)) >> 1) - c.J1 - ((nh0 * (nh0 - 1u)) >> 1);
  const uint32_t lo_term = acc2 - c.n1 * start1 - ((c.n1 * (c.n1 - 1u)) >> 1) - c.n3 * start3 -
                           ((c.n3 * (c.n3 - 1u)) >> 1) - n0 * N0 - c.n2 * (N0 + N2);
  return hi_term + lo_term;
}

// `lead`: the sequence is known to START with that many zero keys (x's first tie group is written in
// ascending y order, so its rows of y's lowest rank -- on left-censored data the rows missing in BOTH
// columns, a fifth of all rows -- come first).  Zero keys are class 0 at every level and the stable
// partition keeps a leading block of them exactly where it is; no one-key ever precedes them, so they
// add nothing to any count.  Threads whose whole range lies inside that block skip both sweeps of
// every level (they only take part in the scans, with zero counts, and in the barriers).
template <bool G>
__device__ __forceinline__ void count_pass(typename Mem<G>::ptr a, typename Mem<G>::ptr b, const int kk,
                                           const int nwarps, const int L, uint32_t* descA, uint32_t* descB,
                                           const int lane, const int warp, const PipeConst pc,
                                           unsigned long long& acc64, const uint32_t lead = 0u) {
  typedef Mem<G> M;
  const uint32_t tid = ((uint32_t)warp << 5) + (uint32_t)lane;
  const uint32_t R = (uint32_t)kk << 3;                // keys per thread range
  const uint32_t my_off = tid * (R << 1);              // bytes
  const uint32_t my_pos = tid * R;
  const uint32_t cap = ((uint32_t)nwarps << 5) * R;    // keys per buffer
  const bool idle = my_pos + R <= lead;
  if (idle && L > 2) {  // the block must read as zeros in the other buffer too (levels alternate)
    const typename M::ptr rb = M::add(b, (int32_t)my_off);
    for (int c = 0; c < kk; ++c) M::st128(M::add(rb, c << 4), 0u, 0u, 0u, 0u);
  }
  for (int s = (L - 1) & ~1; s >= 0; s -= 2) {
    const typename M::ptr ra = M::add(a, (int32_t)my_off);
    uint32_t cl = 0, cb = 0, wm = 0, m4 = 0;
    uint32_t chj[4] = {0u, 0u, 0u, 0u};
#pragma unroll kCountUnroll
    for (int c = 0; c < (idle ? 0 : kk); ++c) {
      uint32_t w[4];
      M::ld128(M::add(ra, c << 4), w[0], w[1], w[2], w[3]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t lo = (w[j] >> s) & 0x00010001u, hi = (w[j] >> (s + 1)) & 0x00010001u;
        cl = M::fadd(cl, lo, pc.one);
        chj[j] = M::fadd(chj[j], hi, pc.one);
        cb = M::fadd(cb, lo & hi, pc.one);
        wm += hi * m4;  // word index 4c + j: the 4c part here, the j part from chj[] below
      }
      m4 += 4u;
    }
    const uint32_t ch = chj[0] + chj[1] + chj[2] + chj[3];
    wm += chj[1] + 2u * chj[2] + 3u * chj[3];
    const LevelCounts lc = fold_counts(cl, ch, cb, wm);
    const uint32_t n1 = lc.n1, n2 = lc.n2, n3 = lc.n3;
    const uint32_t A = n1 | (n2 << 16), B = n3;
    uint32_t inclA = A, inclB = B;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t tA = __shfl_up_sync(FULL, inclA, d), tB = __shfl_up_sync(FULL, inclB, d);
      if (lane >= d) {
        inclA += tA;
        inclB += tB;
      }
    }
    if (lane == 31) {
      descA[warp] = inclA;
      descB[warp] = inclB;
    }
    __syncthreads();
    const uint32_t vA = (lane < nwarps) ? descA[lane] : 0u, vB = (lane < nwarps) ? descB[lane] : 0u;
    const uint32_t totA = __reduce_add_sync(FULL, vA), totB = __reduce_add_sync(FULL, vB);
    const uint32_t exA = __reduce_add_sync(FULL, (lane < warp) ? vA : 0u) + inclA - A;
    const uint32_t exB = __reduce_add_sync(FULL, (lane < warp) ? vB : 0u) + inclB - B;
    const uint32_t e1 = exA & 0xffffu, e2 = exA >> 16, e3 = exB;           // class counts before my range
    const uint32_t N1 = totA & 0xffffu, N2 = totA >> 16, N3 = totB;
    const uint32_t N0 = cap - N1 - N2 - N3;
    const uint32_t start1 = N0 + e1, start3 = N0 + N1 + N2 + e3;
    uint32_t Q01 = (start1 << 16) | (my_pos - e1 - e2 - e3);
    uint32_t Q23 = (start3 << 16) | (N0 + N1 + e2);
    uint32_t acc2 = 0;
    uint32_t bLl = 1u << s, bHl = 2u << s, bLh = 1u << (s + 16), bHh = 2u << (s + 16);
    asm volatile("" : "+r"(bLl), "+r"(bHl), "+r"(bLh), "+r"(bHh));  // keep the bit tests single LOP3s
    if (idle) {
      // nothing to move, nothing to count
    } else if (s > 0) {
#pragma unroll kScatterUnroll
      for (int c = 0; c < kk; ++c) {
        uint32_t w[4];
        M::ld128(M::add(ra, c << 4), w[0], w[1], w[2], w[3]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          M::step4(Q01, Q23, acc2, w[j], bLl, bHl, w[j], b, pc);
#if ICIKT_HI_KEY_MASKS
          M::step4(Q01, Q23, acc2, w[j], bLh, bHh, M::hi16(w[j], pc.c64k), b, pc);
#else
          const uint32_t kh = M::hi16(w[j], pc.c64k);  // the odd key, tested with the even key's masks
          M::step4(Q01, Q23, acc2, kh, bLl, bHl, kh, b, pc);
#endif
        }
      }
    } else {  // the sorted sequence itself is not needed: the last level only counts
#pragma unroll 1
      for (int c = 0; c < kk; ++c) {
        uint32_t w[4];
        M::ld128(M::add(ra, c << 4), w[0], w[1], w[2], w[3]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          M::step4c(Q01, Q23, acc2, w[j], bLl, bHl, pc);
          M::step4c(Q01, Q23, acc2, w[j], bLh, bHh, pc);
        }
      }
    }
    (void)N3;
    if (!idle) acc64 += (unsigned long long)level_share(lc, R, acc2, e2 + e3, start1, start3, N0, N2);
    __syncthreads();  // also protects descA/descB for the next level
    const typename M::ptr t = a;
    a = b;
    b = t;
  }
}

}  // namespace
}  // namespace icikt
