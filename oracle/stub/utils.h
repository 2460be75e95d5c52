/* Stub for the `utils.h` that old/parallel_*.c include but the reference does not
 * ship.  OUR file, test infrastructure only: it supplies the six
 * macros from -D flags so the unmodified old/ sources compile into oracle/_ref. */
#ifndef ORACLE_STUB_UTILS_H
#define ORACLE_STUB_UTILS_H
#ifndef DIM_ROW
#error "define DIM_ROW, DIM_COL, DIM_KERNEL, GBLUR_NORM with -D"
#endif
#define DIM_RGB 3
#define PAD (DIM_KERNEL - 1)
#endif
