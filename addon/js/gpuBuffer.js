// gpuBuffer.js -- src/buffer/bufferCompress.js and src/buffer/bufferDecompress.js re-pointed at the addon.  ensureBuffer and
// the worst-case allocation (bufferCompress.js:135-142) stay in JS; header, block loop and footer (:144-252) are one call.
// Unverified: no JS engine exists in the build image (INTEGRATION.md).
import { ensureBuffer } from '../../src/shared/lz4Util.js';
import { gpu } from './gpuBlock.js';

const EMPTY = new Uint8Array(0);

export function compressBuffer(input, dictionary = null, maxBlockSize = 4194304, blockIndependence = false,
                               contentChecksum = false, addContentSize = true, outputBuffer = null, blockChecksum = false) {
    const raw = ensureBuffer(input);
    const len = raw.length | 0;
    const output = outputBuffer ?? new Uint8Array((19 + len + ((len / 255) | 0) + 64 + 8 + 8 * ((len >> 16) + 1)) | 0);
    const dict = dictionary && dictionary.length > 0 ? ensureBuffer(dictionary) : EMPTY;
    const n = gpu.frameCompress(raw, dict, maxBlockSize >>> 0, !!blockIndependence, !!contentChecksum, !!addContentSize, output,
                                !!blockChecksum);
    // the used portion of the output buffer, whoever allocated it (bufferCompress.js:254-255)
    return output.subarray(0, Math.min(n, output.length));
}

export const decompressBuffer = (input, dictionary = null, verifyChecksum = true) =>
    gpu.frameDecompress(ensureBuffer(input), dictionary ? ensureBuffer(dictionary) : EMPTY, verifyChecksum ? 1 : 0);

/** LZ4Worker.compress / decompress (src/webWorker/workerClient.js:114-152): Promise per task, the work off the JS thread. */
export const compressWorker = (input, o = {}) => {
    const raw = ensureBuffer(input);
    const len = raw.length | 0;
    const output = gpu.allocPinned((19 + len + ((len / 255) | 0) + 64 + 8 + 8 * ((len >> 16) + 1)) | 0);
    const dict = o.dictionary && o.dictionary.length > 0 ? ensureBuffer(o.dictionary) : EMPTY;
    return gpu.frameCompressAsync(raw, dict, (o.maxBlockSize ?? 4194304) >>> 0, !!o.blockIndependence, !!o.contentChecksum,
                                  o.addContentSize ?? true, output, false).then((n) => output.subarray(0, Math.min(n, output.length)));
};
export const decompressWorker = (input, o = {}) =>
    gpu.frameDecompressAsync(ensureBuffer(input), o.dictionary ? ensureBuffer(o.dictionary) : EMPTY, (o.verifyChecksum ?? true) ? 1 : 0);
