/* Shim: lets code written against the reference keep `#include <trico/floating_point_stream_compression.h>`.
 * The declarations live in trico_b200.h (drop-in for /root/reference/trico/floating_point_stream_compression.h). */
#include "../trico_b200.h"
