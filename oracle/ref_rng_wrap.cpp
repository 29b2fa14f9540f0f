// C wrappers around the reference's OWN Random.hpp / Bitmanip.hpp (included from /root/reference at build
// time, never copied). Used by tests/gen_golden.py to produce tests/golden/rng_kat.json and by
// tests/test_oracle_rng.py to cross-check oracle_math.hpp where /root/reference is available.
#include "Random.hpp"
extern "C" {
uint32_t ref_hash_u32(uint32_t i) { return hash_u32(i); }
uint32_t ref_hash_2d(uint32_t x, uint32_t y) { return hash_2d(x, y); }
uint32_t ref_pcg_generate(uint32_t* s) { return pcg_generate(s); }
float ref_rand_unit_float(uint32_t* s) { return rand_unit_float(s); }
uint32_t ref_rand_bounded_int(uint32_t* s, uint32_t range) { return rand_bounded_int(s, range); }
float ref_make_unit_float(uint32_t x) { return make_unit_float(x); }
uint32_t ref_bitreverse(uint32_t x) { return bitreverse(x); }
uint32_t ref_round_up_pow2(uint32_t x) { return round_up_pow2(x); }
}
