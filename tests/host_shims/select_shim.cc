// Host build of slam-sdvl_b200/csrc/select_impl.h so the device corner selector's ordering logic can be checked
// against libstdc++ (through the oracle) without a GPU.
#include "../../slam-sdvl_b200/csrc/select_impl.h"
extern "C" int shim_retain_best22(uint32_t* a, int n, int keep) { return sdvlb_sel::retain_best<22>(a, n, keep); }
extern "C" int shim_retain_best10(uint32_t* a, int n, int keep) { return sdvlb_sel::retain_best<10>(a, n, keep); }
extern "C" void shim_heap_select22(uint32_t* a, int middle, int n) { sdvlb_sel::heap_select<22>(a, middle, n); }
