// gpuBlock.js -- what the bodies of src/block/blockCompress.js, src/block/blockDecompress.js and src/xxhash32/xxhash32.js
// become once the addon is present: the same exports and argument order, every call forwarded to the C ABI.
// Unverified: no JS engine exists in the build image (INTEGRATION.md).
import { createRequire } from 'node:module';

const gpu = createRequire(import.meta.url)('../dlz4.node');          // throws when no CUDA device is usable: no CPU fallback

/** compressBlock(src, output, srcStart, srcLen, hashTable, outputOffset) -> bytes written (src/block/blockCompress.js:31) */
export const compressBlock = (src, output, srcStart, srcLen, hashTable, outputOffset = 0) =>
    gpu.compressBlock(src, output, srcStart | 0, srcLen | 0, hashTable, outputOffset | 0);

/** decompressBlock(input, inputOffset, inputSize, output, outputOffset, dictionary) -> bytes written (blockDecompress.js:30) */
export const decompressBlock = (input, inputOffset, inputSize, output, outputOffset, dictionary) =>
    gpu.decompressBlock(input, inputOffset, inputSize, output, outputOffset, dictionary ?? null);

/** xxHash32(input, seed) -> u32 (src/xxhash32/xxhash32.js:21) */
export const xxHash32 = (input, seed = 0) => gpu.xxHash32(input, seed >>> 0);

export { gpu };
