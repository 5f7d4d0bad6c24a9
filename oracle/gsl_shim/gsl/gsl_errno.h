/* Minimal GSL-compatible shim (TEST INFRASTRUCTURE ONLY).
 * GSL is an external, un-vendored dependency of the reference (FindGSL.cmake);
 * it is absent from this image.  This shim restates the published algorithms of
 * the few GSL entry points the reference calls so that the UNMODIFIED reference
 * sources compile into oracle/_ref/.  Nothing here is shipped in the product. */
#ifndef IS3D_GSL_SHIM_ERRNO_H
#define IS3D_GSL_SHIM_ERRNO_H
#ifdef __cplusplus
extern "C" {
#endif
enum { GSL_SUCCESS = 0, GSL_EDOM = 1, GSL_EINVAL = 4, GSL_ESING = 21 };
typedef void gsl_error_handler_t(const char *reason, const char *file, int line, int gsl_errno);
gsl_error_handler_t *gsl_set_error_handler_off(void);
void gsl_shim_error(const char *reason, const char *file, int line, int gsl_errno);
#ifdef __cplusplus
}
#endif
#endif
