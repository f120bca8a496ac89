/* Shim: lets code written against the reference keep `#include <trico/transpose_aos_to_soa.h>`.
 * The declarations live in trico_b200.h (drop-in for /root/reference/trico/transpose_aos_to_soa.h). */
#include "../trico_b200.h"
