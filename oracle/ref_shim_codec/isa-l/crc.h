// ISA-L is a third-party dependency that is absent here (deps/3rd); ob_crc64.cpp only calls crc32_iscsi from its
// Intel dispatch branch, which the checker never takes (it calls crc64_sse42 / the table versions directly).
#pragma once
#include <stdint.h>
extern "C" unsigned int crc32_iscsi(unsigned char *buffer, int len, unsigned int init_crc);
