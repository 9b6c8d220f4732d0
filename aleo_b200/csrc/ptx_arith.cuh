// ptx_arith.cuh -- 32-bit carry-chain primitives for the sm_100a integer pipes.
//
// ptxas fuses an adjacent `mad(c).lo.cc` / `madc.hi.cc` pair with the same multiplicands and
// neighbouring accumulator registers into ONE `IMAD.WIDE.U32.X` (32x32+64 -> 64 with predicate
// carry in/out), which is what every multiplier row below is shaped to produce (checked with
// cuobjdump -sass; see DESIGN.md "Field arithmetic").
#pragma once
#include "platform.cuh"

namespace ptx {
#ifndef ALEO_EMU
DEV u32 add_cc(u32 a, u32 b) { u32 r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
DEV u32 addc_cc(u32 a, u32 b) { u32 r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
DEV u32 addc(u32 a, u32 b) { u32 r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
DEV u32 sub_cc(u32 a, u32 b) { u32 r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
DEV u32 subc_cc(u32 a, u32 b) { u32 r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
DEV u32 subc(u32 a, u32 b) { u32 r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
DEV u32 mad_lo_cc(u32 a, u32 b, u32 c) { u32 r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
DEV u32 madc_lo_cc(u32 a, u32 b, u32 c) { u32 r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
DEV u32 mad_hi_cc(u32 a, u32 b, u32 c) { u32 r; asm volatile("mad.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
DEV u32 madc_hi_cc(u32 a, u32 b, u32 c) { u32 r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
DEV u32 madc_hi(u32 a, u32 b, u32 c) { u32 r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
DEV u32 mul_lo(u32 a, u32 b) { return a * b; }
// 0 - a, opaque to the optimiser: when NVVM can see that the Montgomery factor is a negation it
// rewrites the following mad pairs and ptxas no longer fuses them into IMAD.WIDE (measured: 2x the
// fma-pipe instructions on the modulus rows).
DEV u32 neg_opaque(u32 a) { u32 r; asm volatile("sub.u32 %0, 0, %1;" : "=r"(r) : "r"(a)); return r; }
// 0 - a, and CF = (a != 0) = the carry of a + (0 - a), i.e. of the limb that a modulus row with unit low limb clears.
// CF comes from an addition (a + 0xffffffff carries iff a != 0): a borrow left by sub.cc is NOT what a following
// addc consumes on sm_100a (measured: wrong products), whatever the PTX manual says about CC.CF.
DEV u32 neg_cc(u32 a) {
  u32 r, t;
  asm volatile("sub.u32 %0, 0, %1;" : "=r"(r) : "r"(a));
  asm volatile("add.cc.u32 %0, %1, 0xffffffff;" : "=r"(t) : "r"(a));
  return r;
}
#else
#define CF (emu::t_cf)
DEV u32 add_cc(u32 a, u32 b) { u64 t = (u64)a + b; CF = (u32)(t >> 32); return (u32)t; }
DEV u32 addc_cc(u32 a, u32 b) { u64 t = (u64)a + b + CF; CF = (u32)(t >> 32); return (u32)t; }
DEV u32 addc(u32 a, u32 b) { return a + b + CF; }
DEV u32 sub_cc(u32 a, u32 b) { u64 t = (u64)a - b; CF = (u32)((t >> 32) & 1); return (u32)t; }      // CF = borrow
DEV u32 subc_cc(u32 a, u32 b) { u64 t = (u64)a - b - CF; CF = (u32)((t >> 32) & 1); return (u32)t; }
DEV u32 subc(u32 a, u32 b) { return a - b - CF; }
DEV u32 mad_lo_cc(u32 a, u32 b, u32 c) { u64 t = (u64)(u32)((u64)a * b) + c; CF = (u32)(t >> 32); return (u32)t; }
DEV u32 madc_lo_cc(u32 a, u32 b, u32 c) { u64 t = (u64)(u32)((u64)a * b) + c + CF; CF = (u32)(t >> 32); return (u32)t; }
DEV u32 mad_hi_cc(u32 a, u32 b, u32 c) { u64 t = (((u64)a * b) >> 32) + c; CF = (u32)(t >> 32); return (u32)t; }
DEV u32 madc_hi_cc(u32 a, u32 b, u32 c) { u64 t = (((u64)a * b) >> 32) + c + CF; CF = (u32)(t >> 32); return (u32)t; }
DEV u32 madc_hi(u32 a, u32 b, u32 c) { return (u32)((((u64)a * b) >> 32) + c + CF); }
DEV u32 mul_lo(u32 a, u32 b) { return a * b; }
DEV u32 neg_opaque(u32 a) { return 0u - a; }
DEV u32 neg_cc(u32 a) { CF = (a != 0u); return 0u - a; }
#undef CF
#endif
}  // namespace ptx
