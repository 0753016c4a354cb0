#include "../../include/passl_b200.h"

static long long g_launches = 0;

extern "C" int passl_b200_version(void) { return 100; }
// kernel-launch counter (host side, not thread safe by design: one Python thread per rank drives the library)
extern "C" long long passl_b200_launch_counter_add(long long n) { g_launches += n; return g_launches; }
extern "C" long long passl_b200_launch_count(void) { return g_launches; }
