/* stb_decode.c -- TEST INFRASTRUCTURE ONLY.
 * Decodes an image with the reference's own vendored stb_image.h (found on the
 * include path at build time, never copied) exactly the way the reference driver
 * does (monolithic/src/main.c:21: stbi_load(path,&w,&h,&ch,0)) and dumps the raw
 * interleaved bytes, so fixtures see the same decoded pixels as the reference. 
 * usage: stb_decode <in.jpg|png> <out.raw>   -> prints "w h channels" */
#define STB_IMAGE_IMPLEMENTATION
#include "stb_image.h"
#include <stdio.h>

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s <in> <out.raw>\n", argv[0]); return 2; }
    int w, h, ch;
    unsigned char *img = stbi_load(argv[1], &w, &h, &ch, 0);
    if (!img) { fprintf(stderr, "decode failed: %s\n", argv[1]); return 1; }
    FILE *f = fopen(argv[2], "wb");
    if (!f) { perror("fopen"); return 1; }
    fwrite(img, 1, (size_t)w * h * ch, f);
    fclose(f);
    printf("%d %d %d\n", w, h, ch);
    stbi_image_free(img);
    return 0;
}
