// Aggregation AMG for K_uu (placeholder until the real hierarchy lands: falls back to block-Jacobi).
#include "common.h"
struct Amg { int dummy; };
void amg_setup(glims_ctx* c) { (void)c; }
void amg_free(glims_ctx* c) { delete c->amg; c->amg = nullptr; }
void amg_vcycle(glims_ctx* c, const double* r, double* z) { (void)c; (void)r; (void)z; }
