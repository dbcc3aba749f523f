// Oobleck VAE entry points (placeholder until the conv kernels land in this round).
#include "../../include/satb200.h"
#include "common.cuh"

using namespace satb;

extern "C" {
int satb_oobleck_create(const SatbOobleckConfig*, SatbOobleck**) { set_last_error("oobleck: not built yet"); return -5; }
void satb_oobleck_destroy(SatbOobleck*) {}
int satb_oobleck_load_weight(SatbOobleck*, const char*, const float*, long long, void*) { set_last_error("oobleck: not built yet"); return -5; }
int satb_oobleck_finalize(SatbOobleck*, void*) { set_last_error("oobleck: not built yet"); return -5; }
int satb_oobleck_decode(SatbOobleck*, const float*, float*, int, int, void*) { set_last_error("oobleck: not built yet"); return -5; }
int satb_oobleck_encode(SatbOobleck*, const float*, float*, int, long long, void*) { set_last_error("oobleck: not built yet"); return -5; }
}
