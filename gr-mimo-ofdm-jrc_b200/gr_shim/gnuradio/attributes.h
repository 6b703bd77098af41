#ifndef JRC_SHIM_GR_ATTRIBUTES_H
#define JRC_SHIM_GR_ATTRIBUTES_H
#define __GR_ATTR_EXPORT __attribute__((visibility("default")))
#define __GR_ATTR_IMPORT __attribute__((visibility("default")))
#define __GR_ATTR_ALIGNED(x) __attribute__((aligned(x)))
#endif
