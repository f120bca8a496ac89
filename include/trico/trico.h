/* Shim: lets code written against the reference keep `#include <trico/trico.h>`.
 * The declarations live in trico_b200.h (drop-in for /root/reference/trico/trico.h). */
#include "../trico_b200.h"
