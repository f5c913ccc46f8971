// Device FLAC decoder of the audio ingest (replaces the per-item librosa decode of
// data_utils_SSL.py:109-113 for the corpus format: 16-bit mono FLAC).  One THREAD per FLAC frame (flac_frame.h: Rice decode +
// predictor restore are sequential inside a frame, frames are independent): a batch of 64 four-second clips is ~1000 frames of
// 4096 samples, i.e. 32 warps spread over 32 SMs for a few hundred microseconds, on a side stream next to the previous batch's
// forward.  The host keeps the cheap, byte-serial part (slsb_flac_scan: frame boundaries, CRC-8, CRC-16 - one pass over the
// bytes at memory speed); the compressed bytes (about half the size of the PCM) are what crosses PCIe.
//
// slsb_flac_decode_frames_host runs the SAME frame core on the CPU: the `-m "not gpu"` tests pin it against flac_decode.cpp and
// the RFC 9639 example files, the GPU tests pin the kernel against it bit for bit.
#include "common.cuh"
#include "kernels.h"
#include "flac_frame.h"

namespace slsb {
namespace {

__global__ void __launch_bounds__(32) flac_frames_kernel(const uint8_t* __restrict__ bytes, const slsb_flac_frame* __restrict__ frames, int n_frames,
                                                         int16_t* __restrict__ pcm, int32_t* __restrict__ status) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_frames) return;
    const slsb_flac_frame f = frames[i];
    status[i] = flacf::decode_frame_mono16(bytes + f.byte_off, f.byte_len, f.bps, f.keep, pcm + f.out_off);
}

}  // namespace

int flac_decode_frames(const uint8_t* bytes_dev, const slsb_flac_frame* frames_dev, int n_frames, int16_t* pcm_dev, int32_t* status_dev, cudaStream_t stream) {
    if (n_frames <= 0) return 0;
    flac_frames_kernel<<<(n_frames + 31) / 32, 32, 0, stream>>>(bytes_dev, frames_dev, n_frames, pcm_dev, status_dev);
    SLSB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

}  // namespace slsb

extern "C" {

int slsb_flac_decode_frames(const uint8_t* bytes_dev, const slsb_flac_frame* frames_dev, int n_frames, int16_t* pcm_dev, int32_t* status_dev, void* stream) {
    if (!bytes_dev || !frames_dev || !pcm_dev || !status_dev) { slsb::set_error("slsb_flac_decode_frames: null buffer"); return -1; }
    return slsb::flac_decode_frames(bytes_dev, frames_dev, n_frames, pcm_dev, status_dev, static_cast<cudaStream_t>(stream));
}

int slsb_flac_decode_frames_host(const uint8_t* bytes, const slsb_flac_frame* frames, int n_frames, int16_t* pcm, int32_t* status) {
    if (!bytes || !frames || !pcm || !status) { slsb::set_error("slsb_flac_decode_frames_host: null buffer"); return -1; }
    for (int i = 0; i < n_frames; ++i)
        status[i] = flacf::decode_frame_mono16(bytes + frames[i].byte_off, frames[i].byte_len, frames[i].bps, frames[i].keep, pcm + frames[i].out_off);
    return 0;
}

}  // extern "C"
