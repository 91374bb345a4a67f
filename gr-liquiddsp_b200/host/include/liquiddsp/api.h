// Export macro of the liquiddsp block library (same name as the reference's
// include/liquiddsp/api.h:26-30 so that downstream code keeps compiling).
#ifndef INCLUDED_LIQUIDDSP_API_H
#define INCLUDED_LIQUIDDSP_API_H
#if defined(_WIN32)
#  define LIQUIDDSP_API __declspec(dllexport)
#elif defined(__GNUC__)
#  define LIQUIDDSP_API __attribute__((visibility("default")))
#else
#  define LIQUIDDSP_API
#endif
#endif
