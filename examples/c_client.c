/* Minimal C client of include/carskit_b200.h: what the JNI glue of INTEGRATION.md does, without the JVM.
 * Trains CAMF_CI (4 factors) for three epochs on a 6-rating toy set and prints the loss of every epoch.
 *
 *   gcc -I include examples/c_client.c -L carskit_b200 -lcarskit_b200 -Wl,-rpath,$PWD/carskit_b200 -o c_client
 *
 * On a box without an sm_100 GPU cars_create() fails with CARS_E_NO_DEVICE (there is no CPU path) and the client
 * says so and exits 3. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "carskit_b200.h"

int main(int argc, char** argv) {
  /* `c_client --gpus N`: ONE handle over N GPUs (cars_desc.num_gpus): users sharded by range inside the library,
   * item block combined with NCCL inside cars_epoch -- the path a single JVM process uses. */
  int gpus = 1;
  for (int a = 1; a + 1 < argc; a++)
    if (strcmp(argv[a], "--gpus") == 0) gpus = atoi(argv[a + 1]);
  /* ratings in the reference's iteration order: user-item pair id ascending, context id ascending */
  const int32_t u[] = {0, 0, 1, 1, 2, 2}, j[] = {0, 1, 0, 1, 0, 1}, ctx[] = {0, 1, 1, 0, 0, 1};
  const double r[] = {4, 5, 3, 4, 2, 5};
  /* two contexts over one dimension with two conditions: context c = condition c */
  const int32_t ctx_ptr[] = {0, 1, 2}, ctx_cond[] = {0, 1};
  enum { U = 3, I = 2, C = 2, F = 4 };
  double P[U * F], Q[I * F], user_bias[U] = {0.01, -0.02, 0.03}, ic_bias[I * C] = {0.5, 0.25, 0.75, 0.125};
  for (int k = 0; k < U * F; k++) P[k] = 0.1 * ((k % 5) - 2);
  for (int k = 0; k < I * F; k++) Q[k] = 0.05 * ((k % 7) - 3);

  cars_desc d;
  memset(&d, 0, sizeof d);
  d.abi_version = CARS_ABI_VERSION;
  d.model = CARS_CAMF_CI;
  d.mode = CARS_EXACT;
  d.schedule = CARS_SCHED_FLAGGED;
  d.num_users = U; d.num_items = I; d.num_conditions = C; d.num_contexts = 2; d.num_factors = F;
  d.nnz = 6; d.u = u; d.j = j; d.ctx = ctx; d.r = r; d.ctx_ptr = ctx_ptr; d.ctx_cond = ctx_cond;
  d.global_mean = 23.0 / 6.0;
  d.reg_u = d.reg_i = d.reg_b = (double)1e-4f; /* Java floats widened to double */
  d.reg_c = (double)1e-3f;
  if (gpus > 1) {
    d.num_gpus = gpus; /* gpu_ids = NULL: devices 0 .. gpus-1 */
    d.combine = CARS_COMBINE_MEAN;
  }

  cars_handle* h = NULL;
  int rc = cars_create(&d, &h);
  if (rc != CARS_OK) {
    printf("cars_create: %d (%s)\n", rc, cars_last_error(NULL));
    return rc == CARS_E_NO_DEVICE ? 3 : 1;
  }
  cars_model_arrays m;
  memset(&m, 0, sizeof m);
  m.P = P; m.Q = Q; m.user_bias = user_bias; m.ic_bias = ic_bias;
  if ((rc = cars_upload(h, &m)) != CARS_OK) { printf("cars_upload: %s\n", cars_last_error(h)); return 1; }
  double lrate = (double)0.02f, loss = 0.0;
  for (int iter = 1; iter <= 3; iter++) {
    if ((rc = cars_epoch(h, lrate, &loss)) != CARS_OK) { printf("cars_epoch: %s\n", cars_last_error(h)); return 1; }
    printf("iter %d: loss = %.17g\n", iter, loss);
  }
  if ((rc = cars_download(h, &m)) != CARS_OK) { printf("cars_download: %s\n", cars_last_error(h)); return 1; }
  cars_stats st;
  cars_get_stats(h, &st);
  printf("P[0][0] = %.17g, gpus = %d, %s\n", P[0], (int)st.num_gpus, cars_version());
  cars_destroy(h);
  return 0;
}
