/* Drop-in copy of the reference's public prototype (monolithic/include/sobel.h:3-5).
 * Implemented by libompimg_b200.so on a B200; dst border bytes are never written. */
#ifndef OMPIMG_DROPIN_SOBEL_H
#define OMPIMG_DROPIN_SOBEL_H
#ifdef __cplusplus
extern "C" {
#endif
void sobel_edge(const unsigned char *src, unsigned char *dst, int width, int height);
#ifdef __cplusplus
}
#endif
#endif
