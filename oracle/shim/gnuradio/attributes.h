/* oracle/shim/gnuradio/attributes.h -- visibility macros used by the
 * reference's include/clenabled/api.h. Test infrastructure only. */
#pragma once
#define __GR_ATTR_EXPORT __attribute__((visibility("default")))
#define __GR_ATTR_IMPORT __attribute__((visibility("default")))
