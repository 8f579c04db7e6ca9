/* oracle/ref_config.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Hand-written stand-in for the config.h that the reference's autotools build would
 * generate (autoconf/automake are not installed here and the reference's own build system
 * is not run).  It is copied to oracle/_ref/config.h and only seen by the reference's
 * sources when oracle/Makefile compiles them in place.  It mirrors the AH_BOTTOM /
 * AC_DEFINE blocks of /root/reference/configure.ac:140-155,205,449-583.
 */
#define HAVE_INLINE 1
#define HAVE_C99_INLINE 1
#define HAVE_EXIT_SUCCESS_AND_FAILURE 1
#define HAVE_DECL_HYPOT 1
#define HAVE_DECL_LOG1P 1
#define HAVE_DECL_EXPM1 1
#define HAVE_DECL_ACOSH 1
#define HAVE_DECL_ASINH 1
#define HAVE_DECL_ATANH 1
#define HAVE_DECL_LDEXP 1
#define HAVE_DECL_FREXP 1
#define HAVE_DECL_ISINF 1
#define HAVE_DECL_ISFINITE 1
#define HAVE_DECL_FINITE 1
#define HAVE_DECL_ISNAN 1
#define HAVE_IEEE_COMPARISONS 1
#define HAVE_IEEE_DENORMALS 1
#define HAVE_GNUX86_IEEE_INTERFACE 1
#define HAVE_PRINTF_LONGDOUBLE 1
#define HAVE_MEMCPY 1
#define HAVE_MEMMOVE 1
#define HAVE_STRDUP 1
#define HAVE_STRTOL 1
#define HAVE_STRTOUL 1
#define HAVE_VPRINTF 1
#define STDC_HEADERS 1
#define PACKAGE "gsl"
#define VERSION "2.4+"
#define GSL_COERCE_DBL(x) (x)
#define DISCARD_POINTER(p) do { ; } while(p ? 0 : 0);
#if defined(GSL_RANGE_CHECK_OFF) || !defined(GSL_RANGE_CHECK)
#define GSL_RANGE_CHECK 0
#endif
#define RETURN_IF_NULL(x) if (!x) { return ; }
#define GSL_DISABLE_DEPRECATED 1
