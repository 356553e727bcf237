// placeholder, replaced below
#include "b7_internal.h"
extern "C" {
int b7_blr_fit(b7_ctx*, const double*, const double*, int, int, const double*, int, b7_blr**, int*) { b7_set_error("blr: not built"); return B7_ERR_STATE; }
int b7_blr_predict(b7_blr*, int, const double*, int64_t, double*, double*) { b7_set_error("blr: not built"); return B7_ERR_STATE; }
int b7_blr_score(b7_blr*, b7_grid*, int, double, int, double, double, double*, int64_t*, int64_t*, double*, int64_t*) { b7_set_error("blr: not built"); return B7_ERR_STATE; }
void b7_blr_free(b7_blr*) {}
}
