#include "../../include/passl_b200.h"
extern "C" int passl_b200_version(void) { return 100; }
