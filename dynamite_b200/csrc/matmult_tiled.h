// Bit-window tiled MatMult for the XOR-structured subspaces (Full->Full and
// same-sector Parity->Parity, with or without XParity).  See matmult_tiled.cu.
#pragma once

#include "context.h"

namespace dnm {

bool tiled_supported(const dnm_mat_s *A);
void tiled_mult(dnm_mat_s *A, dnm_vec_t x, dnm_vec_t y);
void tiled_free(dnm_mat_s *A);
int tiled_passes(dnm_mat_s *A);
int tiled_jit_passes(dnm_mat_s *A);  // passes that run a generated (operator-specialised) kernel
// sharded helpers (row-local, no exchange): diag[local rows], d_out[0] = inf-norm (global)
void tiled_diag(dnm_mat_s *A, double *d_diag);
void tiled_norm(dnm_mat_s *A, double *d_out);

void general_mult(dnm_mat_s *A, const cplx *x, cplx *y);

}  // namespace dnm
