/* Types.h -- the subset of the 7-zip/LZMA-SDK interface types that the libcsc C
 * API is expressed in (reference: src/libcsc/Types.h:137-154 ISeqInStream /
 * ISeqOutStream, :220-225 ICompressProgress, :227-231 ISzAlloc).  Layout- and
 * name-compatible so that callers written against the reference headers
 * (csarc, csc) bind to this library unchanged.  The first vtable argument is the
 * struct pointer itself (callers embed the struct first in their object). */
#ifndef __7Z_TYPES_H
#define __7Z_TYPES_H

#include <stddef.h>
#include <stdint.h>

#ifndef EXTERN_C_BEGIN
#ifdef __cplusplus
#define EXTERN_C_BEGIN extern "C" {
#define EXTERN_C_END }
#else
#define EXTERN_C_BEGIN
#define EXTERN_C_END
#endif
#endif

EXTERN_C_BEGIN

#define SZ_OK 0
#define SZ_ERROR_DATA 1
#define SZ_ERROR_MEM 2
#define SZ_ERROR_READ 8
#define SZ_ERROR_WRITE 9

typedef int SRes;
typedef unsigned char Byte;
typedef short Int16;
typedef unsigned short UInt16;
typedef int Int32;
typedef unsigned int UInt32;
typedef long long int Int64;
typedef unsigned long long int UInt64;
typedef size_t SizeT;
typedef int Bool;
#define True 1
#define False 0

typedef struct
{
  SRes (*Read)(void *p, void *buf, size_t *size);
    /* if (input(*size) != 0 && output(*size) == 0) means end_of_stream.
       (output(*size) < input(*size)) is allowed */
} ISeqInStream;

typedef struct
{
  size_t (*Write)(void *p, const void *buf, size_t size);
    /* Returns: result - the number of actually written bytes.
       (result < size) means error */
} ISeqOutStream;

typedef struct
{
  SRes (*Progress)(void *p, UInt64 inSize, UInt64 outSize);
} ICompressProgress;

typedef struct
{
  void *(*Alloc)(void *p, size_t size);
  void (*Free)(void *p, void *address); /* address can be 0 */
} ISzAlloc;

EXTERN_C_END

#endif
