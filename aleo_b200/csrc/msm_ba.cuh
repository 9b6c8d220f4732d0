// msm_ba.cuh -- batch-affine pair-tree levels in front of the XYZZ bucket accumulation (msm.cuh step 5).
//
// The sorted entry list holds, bucket by bucket, the points a bucket has to sum.  Summing them one after the other
// into an XYZZ accumulator costs 8 products + 2 squares per point (madd-2008-s).  An AFFINE addition costs
// 2 products + 1 square + one inversion, and inversions can be shared (Montgomery's trick: 3 products per
// member and ONE inversion per batch) -- but only between INDEPENDENT additions.  A pair-tree makes them so:
//
//   level l holds, per bucket, ceil(m / 2^l) affine points (level 0 = the gathered bases, sign applied);
//   level l + 1 is built by adding the points of every bucket two by two: positions (s + 2i, s + 2i + 1) of a bucket
//   whose level-l points start at s give output slot s' + i; an odd last point is copied.
//
// Every addition of a level is independent of every other one, so a thread takes a contiguous run of input positions
// (all threads the same length, whatever the bucket sizes are: the flat scheme of the XYZZ kernel), and shares one
// inversion over ALL additions of its run (up to a few hundred):
//   pass 1 (forward)   walks the run, copies single points, and for every pair stores the running product of the
//                      denominators BEFORE it (48 B) and a record {the two points, output slot | doubling flag} (16 B)
//                      in scratch memory (planes of 16-byte words, coalesced across the warp);
//   inversion          one binary-GCD inversion per thread, all lanes of a warp busy;
//   pass 2 (backward)  unwinds: 1 / den_i = inv * prefix_{i-1}, inv *= den_i; lambda = num_i / den_i;
//                      x3 = lambda^2 - x1 - x2, y3 = lambda (x1 - x3) - y1; stores the sum.
// 5 products + 1 square per addition (1590 IMAD.WIDE against 2628) plus 1 / k of an inversion that runs on the ALU
// pipe.  Exceptional pairs never enter the batch: a point at infinity (packed (0, 0)) or P + (-P) is resolved in
// pass 1; P + P joins the batch with den = 2 y, num = 3 x^2.  Level sizes need no per-level bookkeeping beyond an
// exclusive scan of ceil(count / 2^l) (scan_value in msm.cuh).  After L levels the buckets' remaining
// ceil(m / 2^L) points are summed by the XYZZ kernel reading the level array directly (accumulate_kernel<.., DIRECT>).
// The result is the same group element as before, so the normalised output is bit-identical.
//
// Stands in for the same reference function as msm.cuh (snarkvm-algorithms 0.14.5 src/msm/variable_base/batched.rs
// uses batched affine additions on the CPU for the same reason; this is not a port of it -- SURVEY.md 8a row 7).
#pragma once
#include "g1.cuh"

namespace msm {
namespace ba {

constexpr u32 TPB = 128;

struct LevelArgs {
  // level 0 input: bases gathered through the sorted entry list (index | sign << 31)
  const unsigned char* bases;
  u32 stride;
  const u32* sorted;
  // level >= 1 input: packed affine points (96 B, infinity = (0, 0))
  const unsigned char* in;
  const u32* start_in;   // per bucket: first input position
  const u32* cnt0;       // per bucket: level-0 size; the size at level l is ceil(cnt0 / 2^l)
  u32 level;             // l
  const u32* start_out;  // per bucket: first output slot = exclusive scan of ceil(cnt0 / 2^(l + 1))
  unsigned char* out;    // packed affine points of level l + 1
  u32 nb;                // buckets
  u32 nthreads;          // runs the input positions are cut into, at most (a multiple of TPB; the stride of the scratch planes)
  u32 run_target;        // positions per run the host planned for (2 k); with fewer positions than planned (zero digits,
  u32 wave;              // witness-like scalars) fewer threads work, in whole waves of `wave`, so that runs keep that length
  u32 copy_via_l1;       // asynchronous copies allocate in L1 (cp.async.ca) instead of bypassing it (.cg)
  uint4* pre;            // scratch: prefix products, plane (op * 3 + j) * nthreads + thread
  uint4* rec;            // scratch: records {first point, second point, output slot | doubling << 31, -}, op * nthreads + thread;
                         // a point is named by its sorted entry (index | sign << 31) at level 0, by its position above
};

DEV Fq fq_load16(const void* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  Fq r;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const uint4 v = q[i];
    r.l[4 * i] = v.x;
    r.l[4 * i + 1] = v.y;
    r.l[4 * i + 2] = v.z;
    r.l[4 * i + 3] = v.w;
  }
  return r;
}
DEV void fq_store16(void* p, const Fq& v) {
  uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
  for (int i = 0; i < 3; i++) q[i] = make_uint4(v.l[4 * i], v.l[4 * i + 1], v.l[4 * i + 2], v.l[4 * i + 3]);
}

// One input point of a level.  x first (pass 1 needs nothing else for an ordinary pair); y on demand.
// Infinity: level >= 1 and packed bases: (0, 0); Rust-layout bases (stride 104): the flag byte, read when x = 0
// (snarkVM's Affine::zero() is (0, 1, true): an infinity flag never comes with x != 0).
template <bool LEVEL0>
struct PointRef {
  const unsigned char* p;
  u32 neg;
  bool rust;
  // name: the sorted entry (level 0) or the position (above)
  DEV PointRef(const LevelArgs& a, u32 name) {
    if (LEVEL0) {
      p = a.bases + (size_t)(name & 0x7fffffffu) * a.stride;
      neg = name >> 31;
      rust = a.stride >= 97;
    } else {
      p = a.in + (size_t)name * 96;
      neg = 0;
      rust = false;
    }
  }
  DEV static u32 name_of(const LevelArgs& a, u32 pos) { return LEVEL0 ? a.sorted[pos] : pos; }
  DEV Fq x() const { return LEVEL0 ? fq_load8(p) : fq_load16(p); }
  DEV Fq y() const {
    Fq v = LEVEL0 ? fq_load8(p + 48) : fq_load16(p + 48);
    if (LEVEL0 && neg) v = fp_neg(v);
    return v;
  }
  // x is this point's x coordinate (already loaded)
  DEV bool is_infinity(const Fq& xv) const {
    if (!fp_is_zero(xv)) return false;
    if (LEVEL0 && rust) return p[96] != 0;
    return fp_is_zero(LEVEL0 ? fq_load8(p + 48) : fq_load16(p + 48));
  }
};

DEV void store_point(unsigned char* out, u32 slot, const Fq& x, const Fq& y) {
  unsigned char* q = out + (size_t)slot * 96;
  fq_store16(q, x);
  fq_store16(q + 48, y);
}
DEV void store_infinity(unsigned char* out, u32 slot) {
  const Fq z = fp_zero<FqParams>();
  store_point(out, slot, z, z);
}

DEV u32 level_count(const u32* cnt0, u32 g, u32 level) { return (cnt0[g] + ((1u << level) - 1u)) >> level; }

// ---- operand staging -------------------------------------------------------------------------------------------
// A thread's operands (two 96-byte points and a 48-byte prefix per addition) are gathered from all over HBM, and with
// 168 registers per thread an SM holds 12 warps: too few to hide the gathers behind other warps' arithmetic (ncu on the
// first version: long_scoreboard 2.7 warps per issue, heavy pipe at 56 %), and prefetch instructions made it worse.
// So the operands of the NEXT addition travel global -> shared memory asynchronously (cp.async: no registers, no
// warp stall) while the current one is computed from the other half of a double buffer.  Shared layout: 16-byte word j
// of thread t at (j * TPB + t) * 16 (conflict-free LDS.128); words 0-2 x1, 3-5 y1, 6-8 x2, 9-11 y2, 12-14 prefix.
constexpr u32 STAGE_POINT_WORDS = 7;                                  // a 96-byte point that starts 8 bytes into a 16-byte word
constexpr u32 STAGE_WORDS = 2 * STAGE_POINT_WORDS + 3;               // two points + the prefix
constexpr u32 STAGE_BYTES = 2 * STAGE_WORDS * 16 * TPB;              // dynamic shared memory of a CTA (69 632 B; 3 CTAs per SM)

#ifndef ALEO_EMU
// SHIFTED: the points may start 8 bytes into a 16-byte word (caller-owned bases at stride 104 from a 16-byte aligned
// array: every odd index).  The copy then takes the ENCLOSING aligned 16-byte words -- [p - 8, p + 104) is still inside
// the array: the 8 bytes in front belong to the previous point -- and the reader adds the shift.  EXPERIMENT, not the
// default: level 0 is faster with plain loads (msm_host.cuh, where the level kernels are launched, has the figures).
template <bool SHIFTED>
struct AsyncStagerT {
  static constexpr bool ASYNC = true;
  static constexpr bool REREAD = false;
  u32 base;  // shared-memory address of word 0 of this thread, buffer 0
  bool ca;
  DEV AsyncStagerT(unsigned char* smem, u32 via_l1) : ca(via_l1 != 0) { base = (u32)__cvta_generic_to_shared(smem) + threadIdx.x * 16u; }
  DEV u32 word(u32 buf, u32 j) const { return base + (buf * STAGE_WORDS + j) * (16u * TPB); }
  DEV void copy16(u32 dst, const void* src) const {
    if (ca)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    else
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
  }
  // point `which` (0 / 1) of the addition staged in `buf`; with_y = false: the x coordinate only
  DEV void stage_point(u32 buf, u32 which, const unsigned char* p, bool with_y) const {
    const u32 shift = SHIFTED ? (u32)((size_t)p & 8u) : 0u;
    const unsigned char* a = p - shift;
    const u32 n16 = ((with_y ? 96u : 48u) + shift + 15u) >> 4;  // 3 / 4 words for x, 6 / 7 for the point
    const u32 w0 = which * STAGE_POINT_WORDS;
#pragma unroll
    for (u32 k = 0; k < STAGE_POINT_WORDS; k++)
      if (k < n16) copy16(word(buf, w0 + k), a + 16 * k);
  }
  // coordinate at byte offset `off` (0: x, 48: y) of that point
  DEV Fq get_fq(u32 buf, u32 which, const unsigned char* p, u32 off) const {
    Fq r;
    const u32 w0 = which * STAGE_POINT_WORDS;
    if (SHIFTED) {
      const u32 o = (u32)((size_t)p & 8u) + off;
#pragma unroll
      for (u32 i = 0; i < 6; i++) {
        const u32 byte = o + 8 * i;
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(r.l[2 * i]), "=r"(r.l[2 * i + 1]) : "r"(word(buf, w0 + (byte >> 4)) + (byte & 8u)) : "memory");
      }
    } else {
#pragma unroll
      for (u32 k = 0; k < 3; k++)
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.l[4 * k]), "=r"(r.l[4 * k + 1]), "=r"(r.l[4 * k + 2]), "=r"(r.l[4 * k + 3]) : "r"(word(buf, w0 + off / 16 + k)) : "memory");
    }
    return r;
  }
  template <bool LEVEL0>
  DEV void stage_op(u32 buf, const LevelArgs& a, const uint4& r, u32, u32, size_t, bool with_y) const {
    stage_point(buf, 0, PointRef<LEVEL0>(a, r.x).p, with_y);
    stage_point(buf, 1, PointRef<LEVEL0>(a, r.y).p, with_y);
  }
  template <bool LEVEL0>
  DEV Fq get(u32 buf, u32 which, const LevelArgs& a, const uint4& r, u32, u32, size_t, u32 off) const {
    return get_fq(buf, which, PointRef<LEVEL0>(a, which ? r.y : r.x).p, off);
  }
  DEV void advance() const {}
  DEV void end_pass1() const {}
  DEV void commit() const { asm volatile("cp.async.commit_group;" ::: "memory"); }
  DEV void wait_all_but_last() const { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
  // the prefix planes are `plane` 16-byte words apart
  DEV void stage_pre(u32 buf, const uint4* p0, size_t plane) const {
    copy16(word(buf, 2 * STAGE_POINT_WORDS), p0);
    copy16(word(buf, 2 * STAGE_POINT_WORDS + 1), p0 + plane);
    copy16(word(buf, 2 * STAGE_POINT_WORDS + 2), p0 + 2 * plane);
  }
  DEV Fq get_pre(u32 buf, const uint4*, size_t) const {
    Fq r;
#pragma unroll
    for (u32 k = 0; k < 3; k++)
      asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.l[4 * k]), "=r"(r.l[4 * k + 1]), "=r"(r.l[4 * k + 2]), "=r"(r.l[4 * k + 3]) : "r"(word(buf, 2 * STAGE_POINT_WORDS + k)) : "memory");
    return r;
  }
};
typedef AsyncStagerT<false> AsyncStager;         // levels >= 1: packed 96-byte records
typedef AsyncStagerT<true> AsyncStagerShifted;   // level 0
#endif
// No staging: plain loads at the point of use.  The fallback for bases that are not 16-byte aligned, the A/B switch
// (ALEO_B200_MSM_BA_STAGE), and what the emulator runs.
// Pass 1b holds few registers, so there the x coordinates of the NEXT pair are loaded into registers while the current
// product runs (`nxt`, rotated by advance()); pass 2 has no registers to spare and loads at the point of use.
struct DirectStager {
  static constexpr bool ASYNC = false;
  static constexpr bool REREAD = false;
  Fq cur[2], nxt[2];
  bool x_ahead;
  DEV DirectStager(unsigned char*, u32) : x_ahead(true) {}
  DEV void stage_point(u32, u32 which, const unsigned char* p, bool with_y) {
    if (!with_y) nxt[which] = fq_load8(p);
  }
  DEV void advance() {
    cur[0] = nxt[0];
    cur[1] = nxt[1];
  }
  DEV void end_pass1() { x_ahead = false; }
  template <bool LEVEL0>
  DEV void stage_op(u32 buf, const LevelArgs& a, const uint4& r, u32, u32, size_t, bool with_y) {
    stage_point(buf, 0, PointRef<LEVEL0>(a, r.x).p, with_y);
    stage_point(buf, 1, PointRef<LEVEL0>(a, r.y).p, with_y);
  }
  template <bool LEVEL0>
  DEV Fq get(u32 buf, u32 which, const LevelArgs& a, const uint4& r, u32, u32, size_t, u32 off) const {
    return get_fq(buf, which, PointRef<LEVEL0>(a, which ? r.y : r.x).p, off);
  }
  DEV Fq get_fq(u32, u32 which, const unsigned char* p, u32 off) const {
    if (x_ahead && off == 0) return cur[which];
    return fq_load8(p + off);
  }
  DEV void commit() const {}
  DEV void wait_all_but_last() const {}
  DEV void stage_pre(u32, const uint4*, size_t) const {}
  DEV Fq get_pre(u32, const uint4* p0, size_t plane) const {
    const uint4 v0 = p0[0], v1 = p0[plane], v2 = p0[2 * plane];
    Fq pre;
    pre.l[0] = v0.x; pre.l[1] = v0.y; pre.l[2] = v0.z; pre.l[3] = v0.w;
    pre.l[4] = v1.x; pre.l[5] = v1.y; pre.l[6] = v1.z; pre.l[7] = v1.w;
    pre.l[8] = v2.x; pre.l[9] = v2.y; pre.l[10] = v2.z; pre.l[11] = v2.w;
    return pre;
  }
};
// Levels >= 1 at FOUR CTAs per SM.  The level kernel is bound by the dependency stalls of its multiplier chains with 12
// warps per SM (ncu: `wait` 3.4 of 7 warp-cycles per issue, heavy pipe 74 % busy); a fourth CTA needs <= 128 registers
// and <= 56 KB of shared memory.  Registers: pass 2 re-reads its operands from the staging buffer right before each use
// (REREAD) instead of holding x1, y1, x2, y2 and the prefix across five products.  Shared memory: the two points stay
// double-buffered (12 words per buffer), the prefix gets ONE slot that is refilled as soon as the current prefix has
// been read -- 27 words per thread, 55 296 B per CTA.
#ifndef ALEO_EMU
constexpr u32 COMPACT_WORDS = 27;
constexpr u32 COMPACT_BYTES = COMPACT_WORDS * 16 * TPB;
struct CompactStager {
  static constexpr bool ASYNC = true;
  static constexpr bool REREAD = true;
  u32 base;
  DEV CompactStager(unsigned char* smem, u32) { base = (u32)__cvta_generic_to_shared(smem) + threadIdx.x * 16u; }
  DEV u32 word(u32 j) const { return base + j * (16u * TPB); }
  DEV static void copy16(u32 dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
  }
  template <bool LEVEL0>
  DEV void stage_op(u32 buf, const LevelArgs& a, const uint4& r, u32, u32, size_t, bool with_y) const {
    const unsigned char* p1 = PointRef<LEVEL0>(a, r.x).p;
    const unsigned char* p2 = PointRef<LEVEL0>(a, r.y).p;
    const u32 n16 = with_y ? 6u : 3u;
#pragma unroll
    for (u32 k = 0; k < 6; k++) {
      if (k < n16) {
        copy16(word(buf * 12 + k), p1 + 16 * k);
        copy16(word(buf * 12 + 6 + k), p2 + 16 * k);
      }
    }
  }
  DEV Fq ld3(u32 w) const {
    Fq r;
#pragma unroll
    for (u32 k = 0; k < 3; k++)
      asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.l[4 * k]), "=r"(r.l[4 * k + 1]), "=r"(r.l[4 * k + 2]), "=r"(r.l[4 * k + 3]) : "r"(word(w + k)) : "memory");
    return r;
  }
  template <bool LEVEL0>
  DEV Fq get(u32 buf, u32 which, const LevelArgs&, const uint4&, u32, u32, size_t, u32 off) const {
    return ld3(buf * 12 + which * 6 + off / 16);
  }
  DEV void stage_pre(u32, const uint4* p0, size_t plane) const {
    copy16(word(24), p0);
    copy16(word(25), p0 + plane);
    copy16(word(26), p0 + 2 * plane);
  }
  DEV Fq get_pre(u32, const uint4*, size_t) const { return ld3(24); }
  DEV void advance() const {}
  DEV void end_pass1() const {}
  DEV void commit() const { asm volatile("cp.async.commit_group;" ::: "memory"); }
  DEV void wait_all_but_last() const { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
  DEV void wait_all() const { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
};
#endif
// The same re-read order with plain loads: what the emulator runs in place of CompactStager.  (Measured for level 0 at
// four CTAs per SM, 128 registers, no spills: 40.4 against 35.0 ms -- level 0 wants fewer gathers in flight, not more.)
struct DirectRereadStager : DirectStager {
  static constexpr bool REREAD = true;
  DEV DirectRereadStager(unsigned char* smem, u32 f) : DirectStager(smem, f) {}
  DEV void wait_all() const {}
};
#ifdef ALEO_EMU
typedef DirectRereadStager CompactStager;
constexpr u32 COMPACT_BYTES = 0;
#endif
#ifdef ALEO_EMU
typedef DirectStager AsyncStager;  // the emulator has no asynchronous copies
typedef DirectStager AsyncStagerShifted;
#endif

constexpr u32 REC_DOUBLE = 1u << 31;    // record flag: P + P (den = 2 y, num = 3 x^2)
constexpr u32 REC_RESOLVED = 1u << 30;  // record flag: exceptional pair, result already stored by pass 1
constexpr u32 REC_SLOT = REC_RESOLVED - 1u;

// 3 CTAs per SM (168 registers, no spills).  4 (128 registers, ~150 bytes of spills) was measured for level 0, whose gathers
// would like more warps in flight: accumulate 69.8 against 67.3 ms at 2^24.
template <bool LEVEL0, class ST, int MINB = 3>
KERNEL void __launch_bounds__(TPB, MINB) level_kernel(LevelArgs a) {
  DYN_SMEM(unsigned char, smem);
  const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.nthreads) return;  // nthreads is a multiple of the CTA size: whole warps leave
  const u32 first = a.start_in[0];
  const u32 total = a.start_in[a.nb - 1] + level_count(a.cnt0, a.nb - 1, a.level) - first;
  if (total == 0) return;
  u32 nt = a.nthreads;
  {
    const u32 want = (total + a.run_target - 1) / a.run_target;      // threads that give runs of the planned length
    const u32 unit = want >= a.wave ? a.wave : TPB;
    const u64 rounded = ((u64)want + unit - 1) / unit * unit;
    if (rounded < nt) nt = (u32)rounded;
  }
  if (t >= nt) return;  // whole warps again
  const u32 L = (total + nt - 1) / nt;
  // a lane whose run lies beyond the end has nothing to do but stays for the warp's shared inversion
  const bool active = (u64)t * L < total;
  const u32 begin = active ? first + t * L : 0u;
  const u32 end = active ? (((u64)t * L + L < total) ? begin + L : first + total) : 0u;
  const size_t NT = a.nthreads;
  // ---- pass 1a: index walk.  Pairs become records, odd last points of a bucket are copied.
  u32 nops = 0;
  if (active) {
    u32 g;
    {
      u32 lo = 0, hi = a.nb - 1;
      while (lo < hi) {  // bucket that holds position `begin`
        const u32 mid = (lo + hi + 1) >> 1;
        if (a.start_in[mid] <= begin)
          lo = mid;
        else
          hi = mid - 1;
      }
      g = lo;
    }
    u32 s = a.start_in[g], e = s + level_count(a.cnt0, g, a.level);
    u32 so = a.start_out[g];
    u32 pos = begin;
    while (pos < end) {
      while (pos >= e) {  // next non-empty bucket (pos < first + total bounds the walk)
        g++;
        s = a.start_in[g];
        e = s + level_count(a.cnt0, g, a.level);
        so = a.start_out[g];
      }
      const u32 off = pos - s;
      if (off & 1u) {  // second point of a pair that belongs to the run before this one
        pos++;
        continue;
      }
      const u32 slot = so + (off >> 1);
      const u32 name1 = PointRef<LEVEL0>::name_of(a, pos);
      if (pos + 1 >= e) {  // odd last point of its bucket: copied
        const PointRef<LEVEL0> p1(a, name1);
        const Fq x1 = p1.x();
        if (p1.is_infinity(x1))
          store_infinity(a.out, slot);
        else
          store_point(a.out, slot, x1, p1.y());
        pos++;
        continue;
      }
      a.rec[(size_t)nops * NT + t] = make_uint4(name1, PointRef<LEVEL0>::name_of(a, pos + 1), slot, 0u);
      nops++;
      pos += 2;
    }
  }
  ST st(smem, a.copy_via_l1);
  // ---- pass 1b: running product of the denominators; exceptional pairs are resolved here and leave the batch
  Fq prefix = fp_one<FqParams>();
  {
    uint4 r = make_uint4(0, 0, 0, 0), rn = r;
    if (nops) {
      r = a.rec[t];
      st.template stage_op<LEVEL0>(0, a, r, 0, t, NT, false);
      if (nops > 1) rn = a.rec[NT + t];
    }
    st.commit();
    st.advance();
    for (u32 j = 0; j < nops; j++) {
      const u32 buf = j & 1u;
      uint4 rnn = rn;
      if (j + 1 < nops) {
        st.template stage_op<LEVEL0>(buf ^ 1u, a, rn, j + 1, t, NT, false);
        if (j + 2 < nops) rnn = a.rec[(size_t)(j + 2) * NT + t];
      }
      st.commit();
      st.wait_all_but_last();
      const Fq x1 = st.template get<LEVEL0>(buf, 0, a, r, j, t, NT, 0), x2 = st.template get<LEVEL0>(buf, 1, a, r, j, t, NT, 0);
      Fq den = fp_sub(x2, x1);
      u32 flag = 0;
      const bool z1 = fp_is_zero(x1), z2 = fp_is_zero(x2);
      if (z1 || z2 || fp_is_zero(den)) {  // rare: a point at infinity, or the same x twice
        const PointRef<LEVEL0> p1(a, r.x), p2(a, r.y);
        const bool inf1 = p1.is_infinity(x1), inf2 = p2.is_infinity(x2);
        const u32 slot = r.z & REC_SLOT;
        if (inf1 || inf2) {
          flag = REC_RESOLVED;
          if (inf1 && inf2)
            store_infinity(a.out, slot);
          else if (inf1)
            store_point(a.out, slot, x2, p2.y());
          else
            store_point(a.out, slot, x1, p1.y());
        } else if (fp_is_zero(den)) {  // P + P or P + (-P)
          const Fq y1 = p1.y();
          if (!fp_eq(y1, p2.y()) || fp_is_zero(y1)) {
            flag = REC_RESOLVED;
            store_infinity(a.out, slot);
          } else {
            flag = REC_DOUBLE;
            den = fp_dbl(y1);
          }
        }
        if (flag) a.rec[(size_t)j * NT + t] = make_uint4(r.x, r.y, r.z | flag, 0u);
        if (flag & REC_RESOLVED) den = fp_one<FqParams>();
      }
      {
        const size_t o = (size_t)j * 3 * NT + t;
        a.pre[o] = make_uint4(prefix.l[0], prefix.l[1], prefix.l[2], prefix.l[3]);
        a.pre[o + NT] = make_uint4(prefix.l[4], prefix.l[5], prefix.l[6], prefix.l[7]);
        a.pre[o + 2 * NT] = make_uint4(prefix.l[8], prefix.l[9], prefix.l[10], prefix.l[11]);
      }
      prefix = fq_mul_v(prefix, den);
      st.advance();
      r = rn;
      rn = rnn;
    }
  }
  st.end_pass1();
  // ---- ONE inversion per warp.  A binary-GCD inversion is a few ten thousand data-dependent shift / subtract steps:
  // run by 32 lanes at once it diverges into several hundred thousand warp instructions (ncu: it was half of the
  // kernel's instructions even at 256 additions per lane), run by one lane it is a tenth of that.  So the lanes' run
  // products are multiplied up across the warp (prefix and suffix scans over shuffles, 5 products each), lane 0
  // inverts the warp's total, and lane j gets 1 / P_j = (1 / total) * prod_{i < j} P_i * prod_{i > j} P_i.
  Fq inv;
#ifndef ALEO_EMU
  {
    const u32 lane = threadIdx.x & 31u;
    Fq incl = prefix, sfx = prefix;
#pragma unroll 1
    for (u32 d = 1; d < 32; d <<= 1) {
      Fq up, dn;
#pragma unroll
      for (int k = 0; k < FqParams::N; k++) {
        up.l[k] = __shfl_up_sync(0xffffffffu, incl.l[k], d);
        dn.l[k] = __shfl_down_sync(0xffffffffu, sfx.l[k], d);
      }
      const Fq a1 = fq_mul_v(incl, up), a2 = fq_mul_v(sfx, dn);
      if (lane >= d) incl = a1;
      if (lane + d < 32) sfx = a2;
    }
    Fq tot, ex, sx;
#pragma unroll
    for (int k = 0; k < FqParams::N; k++) {
      tot.l[k] = __shfl_sync(0xffffffffu, incl.l[k], 31);
      ex.l[k] = __shfl_up_sync(0xffffffffu, incl.l[k], 1);
      sx.l[k] = __shfl_down_sync(0xffffffffu, sfx.l[k], 1);
    }
    if (lane == 0) ex = fp_one<FqParams>();
    if (lane == 31) sx = fp_one<FqParams>();
    Fq tinv = tot;
    if (lane == 0) tinv = fq_inv_ni(tot);
#pragma unroll
    for (int k = 0; k < FqParams::N; k++) tinv.l[k] = __shfl_sync(0xffffffffu, tinv.l[k], 0);
    inv = fq_mul_v(fq_mul_v(tinv, ex), sx);
  }
#else
  inv = fq_inv_ni(prefix);  // the emulator runs CUDA threads one after the other: every thread inverts its own product
#endif
  if (nops == 0) return;
  // ---- pass 2 (REREAD stagers): the same unwinding with every operand read from the staging buffer right before its use
  if constexpr (ST::REREAD) {
    uint4 r = a.rec[(size_t)(nops - 1) * NT + t], rn = r;
    st.template stage_op<LEVEL0>((nops - 1) & 1u, a, r, nops - 1, t, NT, true);
    st.stage_pre(0, &a.pre[(size_t)(nops - 1) * 3 * NT + t], NT);
    if (nops > 1) rn = a.rec[(size_t)(nops - 2) * NT + t];
    st.commit();
    for (u32 i = nops; i-- > 0;) {
      const u32 buf = i & 1u;
      uint4 rnn = rn;
      if (i > 1) rnn = a.rec[(size_t)(i - 2) * NT + t];
      st.wait_all();
      const bool live = !(r.z & REC_RESOLVED);
      const u32 dbl = r.z >> 31, slot = r.z & REC_SLOT;
      const uint4* pre_i = &a.pre[(size_t)i * 3 * NT + t];
      Fq inv_den;
      if (live) {
        Fq den;
        if (dbl) {
          Fq y1 = st.template get<LEVEL0>(buf, 0, a, r, i, t, NT, 48);
          if (LEVEL0 && (r.x >> 31)) y1 = fp_neg(y1);
          den = fp_dbl(y1);
        } else {
          den = fp_sub(st.template get<LEVEL0>(buf, 1, a, r, i, t, NT, 0), st.template get<LEVEL0>(buf, 0, a, r, i, t, NT, 0));
        }
        inv_den = fq_mul_v(inv, st.get_pre(buf, pre_i, NT));
        inv = fq_mul_v(inv, den);
      }
      if (i) {  // the prefix slot is free again: the next addition's operands set out now
        st.template stage_op<LEVEL0>(buf ^ 1u, a, rn, i - 1, t, NT, true);
        st.stage_pre(0, &a.pre[(size_t)(i - 1) * 3 * NT + t], NT);
      }
      st.commit();
      if (live) {
        Fq y1 = st.template get<LEVEL0>(buf, 0, a, r, i, t, NT, 48);
        if (LEVEL0 && (r.x >> 31)) y1 = fp_neg(y1);
        Fq num;
        if (dbl) {
          const Fq xx = fq_sqr_v(st.template get<LEVEL0>(buf, 0, a, r, i, t, NT, 0));
          num = fp_add(fp_dbl(xx), xx);
        } else {
          Fq y2 = st.template get<LEVEL0>(buf, 1, a, r, i, t, NT, 48);
          if (LEVEL0 && (r.y >> 31)) y2 = fp_neg(y2);
          num = fp_sub(y2, y1);
        }
        const Fq lam = fq_mul_v(num, inv_den);
        const Fq x1 = st.template get<LEVEL0>(buf, 0, a, r, i, t, NT, 0);
        Fq x3 = fp_sub(fq_sqr_v(lam), x1);
        x3 = fp_sub(x3, dbl ? x1 : st.template get<LEVEL0>(buf, 1, a, r, i, t, NT, 0));
        const Fq y3 = fp_sub(fq_mul_v(lam, fp_sub(x1, x3)), y1);
        store_point(a.out, slot, x3, y3);
      }
      r = rn;
      rn = rnn;
    }
    return;
  }
  // ---- pass 2: unwind, last addition first
  {
    uint4 r = a.rec[(size_t)(nops - 1) * NT + t], rn = r;
    st.template stage_op<LEVEL0>((nops - 1) & 1u, a, r, nops - 1, t, NT, true);
    // the prefix planes are NT * 16 bytes apart: staged word by word
    st.stage_pre((nops - 1) & 1u, &a.pre[(size_t)(nops - 1) * 3 * NT + t], NT);
    if (nops > 1) rn = a.rec[(size_t)(nops - 2) * NT + t];
    st.commit();
    for (u32 i = nops; i-- > 0;) {
      const u32 buf = i & 1u;
      uint4 rnn = rn;
      if (i) {
        st.template stage_op<LEVEL0>(buf ^ 1u, a, rn, i - 1, t, NT, true);
        st.stage_pre(buf ^ 1u, &a.pre[(size_t)(i - 1) * 3 * NT + t], NT);
        if (i > 1) rnn = a.rec[(size_t)(i - 2) * NT + t];
      }
      st.commit();
      st.wait_all_but_last();
      if (!(r.z & REC_RESOLVED)) {
        const u32 dbl = r.z >> 31, slot = r.z & REC_SLOT;
        const Fq x1 = st.template get<LEVEL0>(buf, 0, a, r, i, t, NT, 0);
        Fq y1 = st.template get<LEVEL0>(buf, 0, a, r, i, t, NT, 48);
        if (LEVEL0 && (r.x >> 31)) y1 = fp_neg(y1);
        const Fq pre = st.get_pre(buf, &a.pre[(size_t)i * 3 * NT + t], NT);
        Fq x2, num, den;
        if (dbl) {
          x2 = x1;
          den = fp_dbl(y1);
          const Fq xx = fq_sqr_v(x1);
          num = fp_add(fp_dbl(xx), xx);
        } else {
          x2 = st.template get<LEVEL0>(buf, 1, a, r, i, t, NT, 0);
          Fq y2 = st.template get<LEVEL0>(buf, 1, a, r, i, t, NT, 48);
          if (LEVEL0 && (r.y >> 31)) y2 = fp_neg(y2);
          den = fp_sub(x2, x1);
          num = fp_sub(y2, y1);
        }
        const Fq inv_den = fq_mul_v(inv, pre);
        inv = fq_mul_v(inv, den);
        const Fq lam = fq_mul_v(num, inv_den);
        const Fq x3 = fp_sub(fp_sub(fq_sqr_v(lam), x1), x2);
        const Fq y3 = fp_sub(fq_mul_v(lam, fp_sub(x1, x3)), y1);
        store_point(a.out, slot, x3, y3);
      }
      r = rn;
      rn = rnn;
    }
  }
}

}  // namespace ba
}  // namespace msm
