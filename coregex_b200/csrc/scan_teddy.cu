// scan_teddy.cu — the multi-literal engine (Slim / Fat Teddy) on the bitstream kernel's skeleton:
// this translation unit IS scan_bits.cu compiled with CGX_TEDDY (see the header of that file):
// gangs of chunks, per-warp TMA window ring, one look-back per gang, staged coalesced output — with
// per-byte bucket-mask lookups instead of class tests and a per-lane reference loop entered at a safe
// point instead of the marker sweeps.  Exports scan_teddy_chunks / launch_scan_teddy.
#define CGX_TEDDY 1
#ifndef CGX_WARPS
#define CGX_WARPS 19
#endif
#include "scan_bits.cu"
