// pdk.cu -- the one translation unit that carries the persistent decode kernel (pdk.cuh: device code + host entry points)
#define B2_PDK_IMPLEMENTATION
#include "pdk.cuh"
