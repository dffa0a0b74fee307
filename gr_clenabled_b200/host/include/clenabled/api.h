#ifndef INCLUDED_CLENABLED_API_H
#define INCLUDED_CLENABLED_API_H
// symbol visibility of the block library (reference: include/clenabled/api.h:26-30)
#if defined(__GNUC__)
#define CLENABLED_API __attribute__((visibility("default")))
#else
#define CLENABLED_API
#endif
#endif
