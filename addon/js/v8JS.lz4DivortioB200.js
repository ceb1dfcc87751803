// v8JS.lz4DivortioB200.js -- benchmark adapter for the reference's own harness (benchmark/src/libs/shared/baseLib.js:4-45):
// a BaseLib subclass with the call shape of benchmark/src/libs/v8JS/v8JS.lz4Divortio.js:64-95, so that
// benchmark/src/benchWorker.js:47-54 times the GPU path beside the pure-JS one on the same payloads (Silesia or the 219-byte
// record of benchmark/src/base/benchUtils.js:7-22).  bench.py's detail.reference_benchmark row runs the same call shape from
// the Python host.  Unverified: no JS engine exists in the build image (INTEGRATION.md).
import { BaseLib } from '../../benchmark/src/libs/shared/baseLib.js';

export class Lz4DivortioB200 extends BaseLib {
    constructor() {
        super('lz4-divortio-b200', 'divortio-lz4 + dlz4.node', 'NodeJS', 'CUDA');
        this.compressBuffer = null;
        this.decompressBuffer = null;
    }

    async load() {
        const m = await import('./gpuBuffer.js');
        this.compressBuffer = m.compressBuffer;
        this.decompressBuffer = m.decompressBuffer;
        this.lib = m;
    }

    // dict: null, blockSize: 4 MiB, blockIndependence: true, contentChecksum: false, addContentSize: true
    compress(input, outputBuffer) {
        if (!this.compressBuffer) throw new Error('Library not loaded');
        // zero-allocation mode when the harness passes its shared output buffer; either way the frame's subarray comes back
        // (src/buffer/bufferCompress.js:254-255 -- the reference's own adapter still expects a length there and is stale)
        return this.compressBuffer(input, null, 4194304, true, false, true, outputBuffer ?? null);
    }

    decompress(compressedInput, outputBuffer) {
        if (!this.decompressBuffer) throw new Error('Library not loaded');
        return this.decompressBuffer(compressedInput);
    }
}

export default Lz4DivortioB200;
