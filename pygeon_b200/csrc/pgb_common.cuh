// Shared helpers for libpgb (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "pgb.h"

void pgb_set_error(const char* fmt, ...);

#define PGB_CHECK(call)                                                                  \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess) {                                                             \
      pgb_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      return 1;                                                                          \
    }                                                                                    \
  } while (0)

#define PGB_LAUNCH_CHECK() PGB_CHECK(cudaGetLastError())

#define PGB_REQUIRE(cond, msg)                         \
  do {                                                 \
    if (!(cond)) {                                     \
      pgb_set_error("%s:%d: %s", __FILE__, __LINE__, msg); \
      return 2;                                        \
    }                                                  \
  } while (0)

static inline int64_t pgb_cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline unsigned pgb_grid(int64_t n, int block) {
  int64_t g = pgb_cdiv(n, block);
  return (unsigned)(g < 1 ? 1 : g);
}

constexpr int PGB_SMS = 148;  // B200
