/* TEST INFRASTRUCTURE — not product code.
 *
 * Force-included (-include) in front of every reference translation unit when
 * building oracle/_ref.  The reference draws its parent flags from the
 * unseeded libc generator: hem::rand() assembles a 32-bit value from eight
 * `std::rand() % 16` nibbles (reference src/cpp_ext/include/base.hpp:44-50)
 * and rand01() divides by 2^32 (base.hpp:53-56).  To give the reference and
 * the CUDA path the *identical* parent-selection mask without touching the
 * reference sources, `rand` is renamed so that std::rand() resolves to a
 * harness function serving the nibbles of Philox4x32-10(seed, k), k being the
 * running component index.  Defined in oracle/ref_driver.cpp.
 */
#ifndef HEM_ORACLE_REF_RNG_SHIM_H
#define HEM_ORACLE_REF_RNG_SHIM_H
#ifdef __cplusplus
#include <cstdlib>
namespace std { int hemoracle_rand(); }
#define rand hemoracle_rand
#endif
#endif
