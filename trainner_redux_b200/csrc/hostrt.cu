// Host-side runtime pieces of the feed: the side-stream upload step of the prefetcher.
//
// The reference's CUDAPrefetcher (traiNNer/data/prefetch_dataloader.py:418-499) is a dozen Python-level torch calls per
// batch (stream context, wait_stream both ways, one `.to(device, non_blocking=True)` per tensor).  At 0.2 ms of GPU work
// per batch that bookkeeping is a visible share of the step, so the product's prefetcher (prefetch.py) hands the whole
// step to ONE library call: order the copy stream behind the consumer, issue one cudaMemcpyAsync per tensor, record the
// "ready" event.  Events are created / destroyed by the caller through this file too (plain cudaEvent_t handles, timing
// disabled); the library still owns no device state.
#include "otf_common.cuh"

extern "C" int otf_event_create(void** event) {
    using namespace otf;
    OTF_REQUIRE(event != nullptr, OTF_ERR_BAD_ARG, "event_create: null pointer");
    cudaEvent_t ev = nullptr;
    cudaError_t e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    if (e != cudaSuccess) return cuda_fail(e, "cudaEventCreateWithFlags");
    *event = (void*)ev;
    return OTF_OK;
}

extern "C" int otf_event_destroy(void* event) {
    using namespace otf;
    if (event == nullptr) return OTF_OK;
    cudaError_t e = cudaEventDestroy((cudaEvent_t)event);
    if (e != cudaSuccess) return cuda_fail(e, "cudaEventDestroy");
    return OTF_OK;
}

extern "C" int otf_event_query(void* event, int* done) {
    using namespace otf;
    OTF_REQUIRE(event != nullptr && done != nullptr, OTF_ERR_BAD_ARG, "event_query: null pointer");
    cudaError_t e = cudaEventQuery((cudaEvent_t)event);
    if (e == cudaSuccess) {
        *done = 1;
    } else if (e == cudaErrorNotReady) {
        *done = 0;
        (void)cudaGetLastError();  // "not ready" is an answer, not a sticky error
    } else {
        return cuda_fail(e, "cudaEventQuery");
    }
    return OTF_OK;
}

extern "C" int otf_stream_wait_event(void* stream, void* event) {
    using namespace otf;
    OTF_REQUIRE(event != nullptr, OTF_ERR_BAD_ARG, "stream_wait_event: null event");
    cudaError_t e = cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)event, 0);
    if (e != cudaSuccess) return cuda_fail(e, "cudaStreamWaitEvent");
    return OTF_OK;
}

extern "C" int otf_upload_async(int n, void* const* dst_dev, const void* const* src_host, const uint64_t* bytes, void* copy_stream,
                                void* consumer_stream, void* ev_consumed, void* ev_ready) {
    using namespace otf;
    OTF_REQUIRE(n >= 0 && n <= 64, OTF_ERR_BAD_ARG, "upload_async: n must be in [0, 64], got %d", n);
    OTF_REQUIRE(n == 0 || (dst_dev && src_host && bytes), OTF_ERR_BAD_ARG, "upload_async: null table");
    OTF_REQUIRE(ev_ready != nullptr, OTF_ERR_BAD_ARG, "upload_async: null ready event");
    OTF_REQUIRE(copy_stream != consumer_stream, OTF_ERR_BAD_ARG, "upload_async: the copy stream must not be the consumer's stream");
    cudaStream_t cs = (cudaStream_t)copy_stream;
    cudaError_t e;
    if (ev_consumed != nullptr) {
        // the destination slots may still be read by work the consumer has issued so far: the copies go behind it
        if ((e = cudaEventRecord((cudaEvent_t)ev_consumed, (cudaStream_t)consumer_stream)) != cudaSuccess) return cuda_fail(e, "cudaEventRecord");
        if ((e = cudaStreamWaitEvent(cs, (cudaEvent_t)ev_consumed, 0)) != cudaSuccess) return cuda_fail(e, "cudaStreamWaitEvent");
    }
    for (int i = 0; i < n; ++i) {  // one plain asynchronous copy per tensor
        OTF_REQUIRE(dst_dev[i] && src_host[i], OTF_ERR_BAD_ARG, "upload_async: null pointer in entry %d", i);
        if (bytes[i] == 0) continue;
        if ((e = cudaMemcpyAsync(dst_dev[i], src_host[i], (size_t)bytes[i], cudaMemcpyHostToDevice, cs)) != cudaSuccess)
            return cuda_fail(e, "cudaMemcpyAsync");
    }
    if ((e = cudaEventRecord((cudaEvent_t)ev_ready, cs)) != cudaSuccess) return cuda_fail(e, "cudaEventRecord");
    return OTF_OK;
}

extern "C" int otf_download_async(void* dst_host, const void* src_dev, uint64_t bytes, void* copy_stream, void* producer_stream,
                                  void* ev_produced, void* ev_done) {
    using namespace otf;
    OTF_REQUIRE(dst_host && src_dev && ev_produced && ev_done, OTF_ERR_BAD_ARG, "download_async: null pointer");
    OTF_REQUIRE(copy_stream != producer_stream, OTF_ERR_BAD_ARG, "download_async: the copy stream must not be the producer's stream");
    cudaStream_t cs = (cudaStream_t)copy_stream;
    cudaError_t e;
    if ((e = cudaEventRecord((cudaEvent_t)ev_produced, (cudaStream_t)producer_stream)) != cudaSuccess) return cuda_fail(e, "cudaEventRecord");
    if ((e = cudaStreamWaitEvent(cs, (cudaEvent_t)ev_produced, 0)) != cudaSuccess) return cuda_fail(e, "cudaStreamWaitEvent");
    if (bytes != 0 && (e = cudaMemcpyAsync(dst_host, src_dev, (size_t)bytes, cudaMemcpyDeviceToHost, cs)) != cudaSuccess)
        return cuda_fail(e, "cudaMemcpyAsync");
    if ((e = cudaEventRecord((cudaEvent_t)ev_done, cs)) != cudaSuccess) return cuda_fail(e, "cudaEventRecord");
    return OTF_OK;
}

extern "C" int otf_event_synchronize(void* event) {
    using namespace otf;
    OTF_REQUIRE(event != nullptr, OTF_ERR_BAD_ARG, "event_synchronize: null event");
    cudaError_t e = cudaEventSynchronize((cudaEvent_t)event);
    if (e != cudaSuccess) return cuda_fail(e, "cudaEventSynchronize");
    return OTF_OK;
}
