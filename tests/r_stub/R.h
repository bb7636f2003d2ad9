#include <stddef.h>
#include "Rinternals.h"
/* R_ext/Memory.h (included by the real R.h): transient storage reclaimed at the end of the .Call */
char *R_alloc(size_t n, int size);
