/* csc_common.h -- CSCProps and error codes of the libcsc C API
 * (reference: src/libcsc/csc_common.h:11-63).  Field order and types are ABI. */
#ifndef _CSC_COMMON_H_
#define _CSC_COMMON_H_

#include <stdint.h>
#include <stddef.h>

#ifdef _7Z_TYPES_
#include <Types.h>
#else
#include "Types.h"
#endif

#define CSC_PROP_SIZE (4 + 3 + 3)

#define DECODE_ERROR (-96)
#define WRITE_ERROR (-97)
#define READ_ERROR (-98)

static const size_t CSC_WRITE_ABORT = (size_t)-1;

typedef struct _CSCProps {
    size_t dict_size;        /* LZ77 window, 32 KB .. 1 GB */
    uint32_t csc_blocksize;  /* compressed-stream block (RC/BC framing), < 16M */
    uint32_t raw_blocksize;  /* uncompressed bytes per Read()/Write(), < 16M */
    /* encoder only */
    uint8_t hash_bits;       /* 6-byte hash table bits */
    uint8_t hash_width;      /* 6-byte hash table candidates per bucket */
    uint8_t bt_hash_bits;    /* binary-tree head bits (-m5) */
    uint32_t bt_size;        /* binary-tree range, 0 = off */
    uint32_t bt_cyc;         /* binary-tree max steps */
    uint8_t good_len;
    uint8_t lz_mode;         /* 2 lazy parser, 3 price-optimal parser */
    uint8_t DLTFilter;
    uint8_t TXTFilter;
    uint8_t EXEFilter;
} CSCProps;

#endif
