/* Drop-in copy of the reference's public prototype (monolithic/include/parallel_to_grayscale.h:4).
 * Implemented by libompimg_b200.so on a B200; semantics in ompimg_b200.h. */
#ifndef OMPIMG_DROPIN_PARALLEL_TO_GRAYSCALE_H
#define OMPIMG_DROPIN_PARALLEL_TO_GRAYSCALE_H
#ifdef __cplusplus
extern "C" {
#endif
void convert_to_grayscale(unsigned char *data, int width, int height, int channels);
#ifdef __cplusplus
}
#endif
#endif
