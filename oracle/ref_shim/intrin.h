// Minimal stand-in for MSVC's <intrin.h> so that /root/reference/Bitmanip.hpp (included by Random.hpp)
// compiles with g++. Only names that file uses; no arithmetic of the reference is replaced.
#pragma once
#include <x86intrin.h>
#include <climits>
#include <cstdint>
#include <cstddef>
#include <algorithm>
static inline unsigned int __popcnt16(unsigned short v) { return __builtin_popcount(v); }
static inline unsigned int __popcnt(unsigned int v) { return __builtin_popcount(v); }
static inline unsigned long long __popcnt64(unsigned long long v) { return __builtin_popcountll(v); }
static inline unsigned short _byteswap_ushort(unsigned short v) { return __builtin_bswap16(v); }
static inline unsigned long _byteswap_ulong(unsigned long v) { return __builtin_bswap32((unsigned int)v); }
static inline unsigned long long _byteswap_uint64(unsigned long long v) { return __builtin_bswap64(v); }
