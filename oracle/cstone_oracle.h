/* TEST INFRASTRUCTURE ONLY.  Symbol naming of liboracle.so (bound through ctypes in tests/_libs.py):
 *   key-only   : orc_<fn>_u32 | orc_<fn>_u64
 *   key + real : orc_<fn>_u32f | orc_<fn>_u64f | orc_<fn>_u64d
 * See cstone_oracle_impl.h for the functions and the reference file:line each restates. */
#pragma once
