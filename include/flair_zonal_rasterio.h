/* flair_zonal_rasterio.h -- C ABI of libfz_rasterio.so: raster file I/O at the two ends of the zonal path
 * (SURVEY.md section 8(f) rank 1), host code, no CUDA, no GDAL.
 *
 * Replaces, for north-up rasters, what the reference does through rasterio / GDAL:
 *
 *   rasterio.open(path).read(window=..., boundless=True, fill_value=0)   flair_zonal_detection/dataset.py:89-117
 *   rasterio.open(path).bounds / .res / .crs / .profile                 flair_zonal_detection/inference.py:76-132
 *   rasterio.open(out, 'w', compress='lzw', ...).write(..., window=)    flair_zonal_detection/inference.py:157-208,343-352
 *   rio_copy(src, dst, driver='COG', compress='LZW', blocksize=512,
 *            overview_resampling='nearest')                             flair_zonal_detection/postprocess.py:33-52
 *
 * Design: the reference decodes one window per tile on one core and LZW-encodes unaligned windows as they arrive.  Here a
 * raster is read ONCE, block by block on all host cores, straight into the (page-locked) array the GPU upload reads from,
 * and the class raster is encoded ONCE from the array the GPU read-back filled: 512 x 512 blocks, one block per task,
 * compressed in parallel, written in file order.  Classic TIFF and BigTIFF (chosen by size), strips or tiles,
 * uncompressed / LZW / Deflate, horizontal predictor (and, reading, the floating-point predictor of elevation rasters),
 * pixel- or band-interleaved, 8 / 16 / 32-bit samples.
 *
 * Conventions: plain pointers and sizes; every function returns 0 on success or a negative code and leaves a message in
 * fzio_last_error() (thread-local).  Arrays are band-sequential [count][height][width], as rasterio's read() returns them.
 */
#ifndef FLAIR_ZONAL_RASTERIO_H
#define FLAIR_ZONAL_RASTERIO_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FZIO_ABI_VERSION 1

/* compression (the TIFF tag values) */
#define FZIO_COMP_NONE 1
#define FZIO_COMP_LZW 5
#define FZIO_COMP_DEFLATE 8

/* sample formats (the TIFF tag values) */
#define FZIO_FMT_UINT 1
#define FZIO_FMT_INT 2
#define FZIO_FMT_FLOAT 3

/* overview resampling */
#define FZIO_OVR_NEAREST 0 /* GDAL's 'nearest': source pixel floor(0.5 + i * src/dst) (postprocess.py:44) */
#define FZIO_OVR_MODE 1    /* most frequent value of the source footprint, ties -> the smallest value */

int fzio_abi_version(void);
const char* fzio_last_error(void);

/* What rasterio's dataset attributes give the path: size, sample type, block layout, georeferencing.
 * left / top = outer corner of the upper-left pixel (PixelIsArea; a PixelIsPoint file is shifted by half a pixel like
 * GDAL does).  epsg = 0 when the GeoKeys carry no code.  overviews = reduced-resolution IFDs that follow the first. */
typedef struct fzio_info {
    int64_t width, height;
    int32_t count, bits, sample_format, compression, predictor, planar, tiled, block_w, block_h, bigtiff, big_endian;
    int32_t overviews, has_georef, epsg, geographic;
    double left, top, res_x, res_y;
} fzio_info;

/* level 0 = the full-resolution image, k > 0 = the k-th overview */
int fzio_tiff_info(const char* path, int level, fzio_info* out);

/* Boundless windowed read (dataset.py:108-115): rows [row0, row0 + win_h), columns [col0, col0 + win_w) of the listed
 * bands (1-based like rasterio's indexes; bands == NULL reads all), pixels outside the raster = 0.  dst holds
 * n_bands planes of win_h rows; strides in BYTES (dst_row_stride >= win_w * bits / 8).  Samples keep the file's type
 * (bits / sample_format of fzio_tiff_info).  threads <= 0: all host cores. */
int fzio_read_window(const char* path, int level, int64_t row0, int64_t col0, int64_t win_h, int64_t win_w,
                     const int32_t* bands, int32_t n_bands, void* dst, int64_t dst_band_stride, int64_t dst_row_stride,
                     int32_t threads);

/* Options of the writer; zero-initialise, then set what differs.  Defaults (all zero): 512 x 512 tiles, LZW, no
 * predictor, band-interleaved when count > 1, no overviews, plain layout, BigTIFF only when classic offsets cannot
 * address the file, all host cores. */
typedef struct fzio_write_opts {
    int32_t block;            /* tile edge, multiple of 16 (0 -> 512) */
    int32_t compression;      /* FZIO_COMP_* (0 -> LZW, the reference's compress='lzw') */
    int32_t predictor;        /* 1 or 2 (0 -> 1); 2 needs 8-bit samples */
    int32_t deflate_level;    /* 1..9 (0 -> 6) */
    int32_t pixel_interleave; /* 1: PlanarConfiguration = 1 (GDAL INTERLEAVE=PIXEL); 0: one plane per band */
    int32_t overviews;        /* number of factor-2 levels; -1: halve until both sides <= block (GDAL's COG rule) */
    int32_t overview_resampling; /* FZIO_OVR_* */
    int32_t cog;              /* 1: COG layout -- IFDs before data, overview data before full-resolution data,
                                 GDAL's structural-metadata ghost area, 4-byte size leader / trailer per block */
    int32_t bigtiff;          /* 0 auto, 1 force BigTIFF, -1 force classic (fails when too large) */
    int32_t threads;          /* <= 0: all host cores */
    int32_t sample_format;    /* FZIO_FMT_* (0 -> UINT) */
    int32_t bits;             /* 8, 16 or 32 (0 -> 8) */
    int32_t epsg;             /* 0: no CRS key */
    int32_t geographic;       /* 1: GeographicTypeGeoKey, 0: ProjectedCSTypeGeoKey */
    int32_t has_georef;       /* 1: write ModelPixelScale / ModelTiepoint / GeoKeyDirectory from left, top, res */
    int32_t reserved;
    double left, top, res;
} fzio_write_opts;

/* data: [count][height][width] samples of opts->bits, band stride / row stride in BYTES (0 -> dense). */
int fzio_write_geotiff(const char* path, const void* data, int32_t count, int64_t height, int64_t width,
                       int64_t band_stride, int64_t row_stride, const fzio_write_opts* opts);

/* postprocess.py:33-52 convert_to_cog: src (any TIFF the reader handles) -> COG with LZW, 512 blocks, nearest overviews,
 * georeferencing carried over.  The caller removes src (as the reference does) after a 0 return. */
int fzio_convert_to_cog(const char* src_path, const char* dst_path, int32_t threads);

/* The codecs by themselves (tests pin them against libtiff through Pillow).  Return the number of bytes produced, or
 * a negative code.  dst_cap for the encoder: fzio_lzw_bound(n). */
int64_t fzio_lzw_bound(int64_t n);
int64_t fzio_lzw_encode(const uint8_t* src, int64_t n, uint8_t* dst, int64_t dst_cap);
int64_t fzio_lzw_decode(const uint8_t* src, int64_t n, uint8_t* dst, int64_t dst_cap);

#ifdef __cplusplus
}
#endif
#endif
