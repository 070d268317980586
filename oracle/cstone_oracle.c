/* TEST INFRASTRUCTURE ONLY — see cstone_oracle_impl.h.  Instantiates the oracle for
 * keys {u32,u64} and (key,real) in {(u32,float),(u64,float),(u64,double)}.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load the resulting liboracle.so. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---------------- 32-bit keys ---------------- */
#define KEY uint32_t
#define KBITS 32
#define MAXLEVEL 10
#define UNUSED_BITS 2
#define KS(name) name##_u32
#define ORC_EMIT_KEY
#include "cstone_oracle_impl.h"
#undef ORC_EMIT_KEY

#define ORC_EMIT_KEY_REAL
#define REAL float
#define RFLOOR floorf
#define RCEIL ceilf
#define RRINT rintf
#define RFABS fabsf
#define KT(name) name##_u32f
#include "cstone_oracle_impl.h"
#undef REAL
#undef RFLOOR
#undef RCEIL
#undef RRINT
#undef RFABS
#undef KT
#undef ORC_EMIT_KEY_REAL
#undef NODE_RANGE0
#undef KEY
#undef KBITS
#undef MAXLEVEL
#undef UNUSED_BITS
#undef KS

/* ---------------- 64-bit keys ---------------- */
#define KEY uint64_t
#define KBITS 64
#define MAXLEVEL 21
#define UNUSED_BITS 1
#define KS(name) name##_u64
#define ORC_EMIT_KEY
#include "cstone_oracle_impl.h"
#undef ORC_EMIT_KEY

#define ORC_EMIT_KEY_REAL
#define REAL float
#define RFLOOR floorf
#define RCEIL ceilf
#define RRINT rintf
#define RFABS fabsf
#define KT(name) name##_u64f
#include "cstone_oracle_impl.h"
#undef REAL
#undef RFLOOR
#undef RCEIL
#undef RRINT
#undef RFABS
#undef KT

#define REAL double
#define RFLOOR floor
#define RCEIL ceil
#define RRINT rint
#define RFABS fabs
#define KT(name) name##_u64d
#include "cstone_oracle_impl.h"
