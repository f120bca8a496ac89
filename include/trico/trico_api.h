/* Shim for /root/reference/trico/trico_api.h: TRICO_API marks exported symbols. */
#ifndef TRICO_API
#define TRICO_API __attribute__((visibility("default")))
#endif
