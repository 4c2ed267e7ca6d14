/* examples/solve_qcqp.c -- the C ABI without any Python: a batch of small non-convex QCQPs
 *     min 0.5 x'Qx + q'x   s.t.  0.5 x'P_j x + a_j'x <= b_j   (j < m)
 * described once as a structure (one QUADFORM block of LEq rows, everything per problem) and solved with
 * sco_solve_batch_host (host buffers in, host buffers out).
 *
 *   gcc -O2 -I include examples/solve_qcqp.c -o examples/solve_qcqp sco_py_b200/libsco_b200.so -lm \
 *       -Wl,-rpath,'$ORIGIN/../sco_py_b200'
 *   ./examples/solve_qcqp 64
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sco_b200.h"

static double urand(unsigned *s) { *s = *s * 1664525u + 1013904223u; return (double)(*s >> 8) / (1u << 24); }
static double nrand(unsigned *s) { return sqrt(-2.0 * log(urand(s) + 1e-12)) * cos(6.283185307179586 * urand(s)); }

int main(int argc, char **argv) {
  const int B = argc > 1 ? atoi(argv[1]) : 16, n = 8, m = 6, ntri = n * (n + 1) / 2;
  /* parameter row: Q (n*n) | q (n) | P packed upper (m*ntri) | a (m*n) | b (m) */
  const long oQ = 0, oq = oQ + n * n, oP = oq + n, oa = oP + (long)m * ntri, ob = oa + (long)m * n, stride = ob + m;
  sco_structure_desc d;
  memset(&d, 0, sizeof(d));
  d.n = n; d.m_lin = 0; d.n_blocks = 1; d.n_groups = 1; d.stride = stride; d.shared_len = 0;
  d.Q.off = oQ; d.q.off = oq; d.c.off = -1; d.lin_l.off = -1; d.lin_u.off = -1; d.obj_prog.off = -1;
  d.blocks[0].family = SCO_FAM_QUADFORM; d.blocks[0].cnt_type = SCO_CNT_LEQ; d.blocks[0].m = m;
  d.blocks[0].group_mask = 1; d.blocks[0].jw = n; d.blocks[0].ipar[0] = n; d.blocks[0].ipar[1] = m;
  d.blocks[0].par.off = oP; d.blocks[0].val.off = ob;

  double *params = calloc((size_t)B * stride, sizeof(double)), *x0 = malloc(sizeof(double) * B * n);
  double *x = malloc(sizeof(double) * B * n), *vio = malloc(sizeof(double) * B), *obj = malloc(sizeof(double) * B);
  int32_t *verdict = malloc(sizeof(int32_t) * B), *stats = malloc(sizeof(int32_t) * 4 * B);
  unsigned seed = 7;
  for (int b = 0; b < B; b++) {
    double *row = params + (size_t)b * stride, M[64];
    for (int i = 0; i < n * n; i++) M[i] = nrand(&seed);
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) {
        double s = 0.0;
        for (int k = 0; k < n; k++) s += M[k * n + i] * M[k * n + j];
        row[oQ + i * n + j] = s / n + (i == j ? 0.1 : 0.0);
      }
    for (int i = 0; i < n; i++) row[oq + i] = nrand(&seed);
    for (int j = 0; j < m; j++) {
      int e = 0;
      for (int r = 0; r < n; r++)
        for (int c = r; c < n; c++) row[oP + (long)j * ntri + e++] = nrand(&seed) / sqrt((double)n); /* indefinite */
      for (int i = 0; i < n; i++) row[oa + (long)j * n + i] = nrand(&seed);
      row[ob + j] = 0.5 + urand(&seed);
    }
    for (int i = 0; i < n; i++) x0[b * n + i] = nrand(&seed);
  }

  sco_handle *h = NULL;
  if (sco_create(&d, 0, &h) != SCO_OK) { fprintf(stderr, "sco_create: %s\n", sco_last_error()); return 1; }
  sco_settings s;
  sco_default_settings(&s);
  s.initial_penalty_coeff = 1.0;       /* tests/sco_osqp/test_solver.py:15-25 */
  s.max_merit_coeff_increases = 5;
  s.min_trust_region_size = 1e-5;
  if (sco_solve_batch_host(h, B, params, x0, &s, x, verdict, NULL, obj, vio, stats) != SCO_OK) {
    fprintf(stderr, "sco_solve_batch_host: %s\n", sco_last_error());
    return 1;
  }
  int conv = 0;
  double worst = 0.0;
  long iters = 0;
  for (int b = 0; b < B; b++) {
    conv += verdict[b] == SCO_VERDICT_CONVERGED;
    if (verdict[b] == SCO_VERDICT_CONVERGED && vio[b] > worst) worst = vio[b];
    iters += stats[4 * b + 2];
  }
  printf("%d problems, %d converged, largest violation among them %.2e, %.0f ADMM iterations per problem\n", B, conv,
         worst, (double)iters / B);
  sco_destroy(h);
  return conv > 0 && worst <= 1e-4 ? 0 : 2;
}
