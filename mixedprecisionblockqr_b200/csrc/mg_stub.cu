// Placeholder definitions until mg.cu / tsqr.cu land (keeps every symbol of include/mpqr.h exported).
#include "common.cuh"
using namespace mpqr;
extern "C" {
int mpqr_mg_get_unique_id(void*) { set_error("multi-GPU path not built yet"); return MPQR_ESTATE; }
int mpqr_mg_create(mpqr_handle**, int, int, int, int, unsigned, int, int, const void*) { set_error("multi-GPU path not built yet"); return MPQR_ESTATE; }
int mpqr_mg_local_cols(const mpqr_handle*) { return MPQR_ESTATE; }
int mpqr_mg_global_col(const mpqr_handle*, int) { return MPQR_ESTATE; }
int mpqr_mg_factor_device(mpqr_handle*, float*, long, void*) { set_error("multi-GPU path not built yet"); return MPQR_ESTATE; }
int mpqr_tsqr_device(const float*, long, long, int, float*, long, float*, long, void*) { set_error("TSQR path not built yet"); return MPQR_ESTATE; }
}
