// temporary stubs (replaced by kl_hh.cu / kl_bicgstab.cu / kl_lanczos.cu)
#include "kl_internal.cuh"
extern "C" {
int kl_gmres_hh_omp(kl_handle_t h, const kl_operator_t *, const double *, double *, int, int, int, double, double *, double *, int *, int *) { return h ? h->fail(KL_ERR_UNSUPPORTED, "not built yet") : KL_ERR_INVALID; }
int kl_gmres_hh_prec_omp(kl_handle_t h, const kl_operator_t *, const double *, double *, int, int, int, double, double *, double *, int *, int *, const kl_precond_t *, const double *, int) { return h ? h->fail(KL_ERR_UNSUPPORTED, "not built yet") : KL_ERR_INVALID; }
int kl_bicgstab(kl_handle_t h, const kl_operator_t *, const double *, double *, int, int, double, int *, double *) { return h ? h->fail(KL_ERR_UNSUPPORTED, "not built yet") : KL_ERR_INVALID; }
int kl_pbicgstab(kl_handle_t h, const kl_operator_t *, const double *, double *, int, int, double, int *, double *, const kl_precond_t *, const double *, int) { return h ? h->fail(KL_ERR_UNSUPPORTED, "not built yet") : KL_ERR_INVALID; }
int kl_pbicgstab_omp(kl_handle_t h, const kl_operator_t *, const double *, double *, int, int, double, int *, double *, const kl_precond_t *, const double *, int) { return h ? h->fail(KL_ERR_UNSUPPORTED, "not built yet") : KL_ERR_INVALID; }
int kl_lanczos(kl_handle_t h, const kl_operator_t *, int, int, int, double *, double *) { return h ? h->fail(KL_ERR_UNSUPPORTED, "not built yet") : KL_ERR_INVALID; }
}
