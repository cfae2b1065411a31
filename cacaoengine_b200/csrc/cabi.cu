// cabi.cu -- the extern "C" layer declared in include/cacao_img.h: device contexts, the
// device-buffer / upload / readback layer, and the host-memory pipelines around the kernels.
//
// There is deliberately no CPU implementation of any pixel loop in this file or anywhere in the
// library: without a CUDA device every compute entry point returns CACAO_E_NO_DEVICE.
#include <cuda.h> // CUmem* types for the shareable allocations; entry points come from cudaGetDriverEntryPoint
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <map>
#include <chrono>
#include <condition_variable>
#include <thread>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include "cacao_img.h"
#include "convert_dispatch.cuh"
#include "equirect.cuh"
#include "mip.cuh"
#include "misc_kernels.cuh"

namespace {

	using namespace cacao;

	// ---- error plumbing ------------------------------------------------------------------------
	thread_local std::string t_last_error;

	int fail(int status, const char* fmt, ...) {
		char buf[512];
		va_list ap;
		va_start(ap, fmt);
		vsnprintf(buf, sizeof buf, fmt, ap);
		va_end(ap);
		t_last_error = buf;
		return status;
	}

	// CACAO_TRACE=1: one line on stderr per host-memory call -- what ran, how many bytes went each way,
	// which path it took and how long it blocked the caller.  (Device-memory calls only enqueue; time
	// those with CUDA events on your stream, or with nsys / ncu.)
	struct HostTrace {
		const char* what;
		const char* path = "";
		uint64_t in_bytes = 0, out_bytes = 0;
		std::chrono::steady_clock::time_point t0;
		bool on;
		explicit HostTrace(const char* w) : what(w), t0(std::chrono::steady_clock::now()) {
			static const bool enabled = getenv("CACAO_TRACE") && getenv("CACAO_TRACE")[0] == '1';
			on = enabled;
		}
		~HostTrace() {
			if(!on) return;
			const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
			std::fprintf(stderr, "[cacao_img] %-18s %-22s up %10llu B  down %10llu B  %9.3f ms  (%.2f GB/s through the call)\n", what, path, static_cast<unsigned long long>(in_bytes), static_cast<unsigned long long>(out_bytes), ms,
						 ms > 0 ? static_cast<double>(in_bytes + out_bytes) / ms / 1e6 : 0.0);
		}
	};

	int fail_cuda(cudaError_t e, const char* what) {
		// "no device" style failures get their own status so callers can tell them from kernel faults
		const bool nodev = (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorInvalidDevice || e == cudaErrorInitializationError);
		return fail(nodev ? CACAO_E_NO_DEVICE : CACAO_E_CUDA, "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
	}

#define CU_TRY(expr)                                         \
	do {                                                     \
		cudaError_t e__ = (expr);                            \
		if(e__ != cudaSuccess) return fail_cuda(e__, #expr); \
	} while(0)

	// ---- the 16->8 table -----------------------------------------------------------------------
	const uint16_t kLutThresholds[254] = {
#include "lut_thresholds.inc"
	};

	struct HostLut {
		uint8_t table[65536];
		uint16_t two_level[4096];             // see lut_16to8<kLutTwoLevel> in convert_pixel.cuh
		uint32_t lanes[cacao::kLanesEntries]; // see lut_16to8<kLutLanes>
		bool two_level_ok = true;
		bool lanes_ok = true;
		HostLut() {
			// level k starts at kLutThresholds[k-1]; the table is monotone with unit steps
			int level = 0;
			for(uint32_t i = 0; i < 65536u; ++i) {
				while(level < 254 && i >= kLutThresholds[level]) ++level;
				table[i] = static_cast<uint8_t>(level);
			}
			for(uint32_t k = 0; k < 4096u; ++k) {
				const uint32_t base = table[16 * k];
				uint32_t step = 16;
				for(uint32_t o = 1; o < 16; ++o)
					if(table[16 * k + o] != base) {
						step = o;
						break;
					}
				two_level[k] = static_cast<uint16_t>(base | (step << 8));
				for(uint32_t o = 0; o < 16; ++o) // the compact form must reproduce every entry
					if(base + (o >= step ? 1u : 0u) != table[16 * k + o]) two_level_ok = false;
			}
			// log-indexed form: segment = upper half-word of float(v + bias), position = lower half-word
			auto float_bits = [](uint32_t v) {
				const float f = static_cast<float>(v + static_cast<uint32_t>(cacao::kLanesBias));
				uint32_t b;
				std::memcpy(&b, &f, 4);
				return b;
			};
			bool seen[cacao::kLanesEntries] = {};
			for(uint32_t v = 0; v < 65536u; ++v) {
				const uint32_t b = float_bits(v), seg = (b >> 16) - cacao::kLanesHiBase;
				if(seg >= static_cast<uint32_t>(cacao::kLanesEntries)) {
					lanes_ok = false;
					break;
				}
				if(seen[seg]) continue;
				seen[seg] = true;
				// v is the first input of its segment: find the one step inside it, if any
				uint32_t step_pos = 0x10000u;
				for(uint32_t u = v + 1; u < 65536u && (float_bits(u) >> 16) == (b >> 16); ++u)
					if(table[u] != table[v]) {
						step_pos = float_bits(u) & 0xFFFFu;
						break;
					}
				lanes[seg] = (static_cast<uint32_t>(table[v]) << 16) + (0x10000u - step_pos) - ((b >> 16) << 16);
			}
			for(int k = 0; k < cacao::kLanesEntries; ++k) lanes_ok = lanes_ok && seen[k];
			for(uint32_t v = 0; v < 65536u && lanes_ok; ++v) { // the form must reproduce every entry
				const uint32_t b = float_bits(v), sum = b + lanes[(b >> 16) - cacao::kLanesHiBase];
				if((sum >> 16) != table[v]) lanes_ok = false;
			}
		}
	};
	const HostLut& host_lut() {
		static const HostLut lut;
		return lut;
	}

	// ---- host-side copy pool --------------------------------------------------------------------
	// The reference API hands over pixels in std::vector memory, i.e. pageable.  A cudaMemcpyAsync from
	// pageable memory is staged by the driver on one thread (measured on the B200 box: 1.3 Gpixel/s end
	// to end for config 2 against 6.6 from pinned memory).  Instead, pageable buffers are copied into /
	// out of a small pinned ring by several host threads (15 GB/s per thread, 50-70 GB/s with 8-16 on
	// the box) while the DMA engines and the kernel work on neighbouring chunks.
	// memcpy for the staging copies: the destination is not read again by this core (it is DMA'd away,
	// or it is the caller's result buffer), so the body is written with non-temporal stores -- no
	// read-for-ownership of the destination lines, a third less DRAM traffic than a cached copy of a
	// slice too small for libc's own streaming threshold.  CACAO_COPY_NT=0 switches back to memcpy.
	inline void stream_copy(char* dst, const char* src, size_t n) {
#if defined(__SSE2__)
		static const bool nt = !(getenv("CACAO_COPY_NT") && getenv("CACAO_COPY_NT")[0] == '0');
		if(nt && n >= 4096) {
			const size_t head = (64 - (reinterpret_cast<uintptr_t>(dst) & 63u)) & 63u;
			if(head) {
				std::memcpy(dst, src, head);
				dst += head;
				src += head;
				n -= head;
			}
			const size_t body = n & ~static_cast<size_t>(63);
			for(size_t o = 0; o < body; o += 64) {
				const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + o));
				const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + o + 16));
				const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + o + 32));
				const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + o + 48));
				_mm_stream_si128(reinterpret_cast<__m128i*>(dst + o), a);
				_mm_stream_si128(reinterpret_cast<__m128i*>(dst + o + 16), b);
				_mm_stream_si128(reinterpret_cast<__m128i*>(dst + o + 32), c);
				_mm_stream_si128(reinterpret_cast<__m128i*>(dst + o + 48), d);
			}
			_mm_sfence();
			if(n > body) std::memcpy(dst + body, src + body, n - body);
			return;
		}
#endif
		std::memcpy(dst, src, n);
	}

	class CopyPool {
	  public:
		static CopyPool& get() {
			static CopyPool pool;
			return pool;
		}
		// blocking parallel memcpy; the caller works too.  About one thread per MiB and no more: all
		// workers are woken (one notify_all is the fastest way to start many), but only the first
		// `helpers` of them touch the job -- fifteen threads fighting over a 6 MB chunk cost more than they
		// saved (2048^2 RGBA16 -> RGBA8 from pageable memory: 4.6 ms with every worker in, 2.5 ms with three).
		void copy(void* dst, const void* src, size_t n) {
			constexpr size_t kPerThread = 1u << 20;
			const size_t parts = std::min(workers_.size() + 1, n / kPerThread);
			if(parts <= 1) {
				stream_copy(static_cast<char*>(dst), static_cast<const char*>(src), n);
				return;
			}
			std::lock_guard<std::mutex> one_at_a_time(call_mutex_);
			const size_t slice = ((n / parts) + 63) & ~static_cast<size_t>(63);
			{
				std::lock_guard<std::mutex> lock(m_);
				dst_ = static_cast<char*>(dst);
				src_ = static_cast<const char*>(src);
				total_ = n;
				slice_ = slice;
				next_ = 0;
				pending_ = (n + slice - 1) / slice;
				helpers_ = parts - 1;
				++epoch_;
			}
			cv_.notify_all();
			work();
			std::unique_lock<std::mutex> lock(m_);
			done_cv_.wait(lock, [this] { return pending_ == 0; });
		}

	  private:
		CopyPool() {
			unsigned hw = std::thread::hardware_concurrency();
			unsigned n = hw > 2 ? std::min(hw - 1, 15u) : 0;
			if(const char* e = getenv("CACAO_COPY_THREADS")) n = static_cast<unsigned>(std::max(0, atoi(e)));
			for(unsigned i = 0; i < n; ++i) workers_.emplace_back([this, i] { loop(i); });
		}
		~CopyPool() {
			{
				std::lock_guard<std::mutex> lock(m_);
				stop_ = true;
			}
			cv_.notify_all();
			for(auto& t : workers_) t.join();
		}
		void work() {
			for(;;) {
				size_t off;
				{
					std::lock_guard<std::mutex> lock(m_);
					if(next_ * slice_ >= total_) return;
					off = next_++ * slice_;
				}
				stream_copy(dst_ + off, src_ + off, std::min(slice_, total_ - off));
				std::lock_guard<std::mutex> lock(m_);
				if(--pending_ == 0) done_cv_.notify_all();
			}
		}
		void loop(size_t index) {
			uint64_t seen = 0;
			for(;;) {
				{
					std::unique_lock<std::mutex> lock(m_);
					cv_.wait(lock, [&] { return stop_ || epoch_ != seen; });
					if(stop_) return;
					seen = epoch_;
					if(index >= helpers_) continue;
				}
				work();
			}
		}
		std::vector<std::thread> workers_;
		std::mutex m_, call_mutex_;
		std::condition_variable cv_, done_cv_;
		char* dst_ = nullptr;
		const char* src_ = nullptr;
		size_t total_ = 0, slice_ = 1, next_ = 0, pending_ = 0, helpers_ = 0;
		uint64_t epoch_ = 0;
		bool stop_ = false;
	};

	// true when `p` is ordinary pageable host memory (not cudaHostAlloc'ed / registered)
	bool is_pageable(const void* p) {
		cudaPointerAttributes attr;
		if(cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
			cudaGetLastError();
			return true;
		}
		return attr.type == cudaMemoryTypeUnregistered;
	}

	// ---- per-device context --------------------------------------------------------------------
	constexpr int kMaxDevices = 64;
	constexpr int kSlots = 3; // staging ring depth of the host-memory pipeline

	struct DeviceCtx {
		int device = -1;
		int sm_count = 0;
		size_t smem_optin = 0;
		uint8_t* lut_dev = nullptr;
		uint16_t* lut2_dev = nullptr;
		uint32_t* lanes_dev = nullptr;
		cudaStream_t s_in = nullptr, s_run = nullptr, s_out = nullptr;
		cudaEvent_t ev_in[kSlots] = {}, ev_run[kSlots] = {}, ev_out[kSlots] = {};
		uint8_t* stage_in[kSlots] = {};
		uint8_t* stage_out[kSlots] = {};
		size_t stage_in_cap[kSlots] = {}, stage_out_cap[kSlots] = {};
		uint8_t* pin_in[kSlots] = {}; // pinned ring for pageable host buffers
		uint8_t* pin_out[kSlots] = {};
		size_t pin_in_cap = 0, pin_out_cap = 0;
		unsigned long long* checksum_dev = nullptr;
		std::mutex pipe_mutex; // one host-memory pipeline call per device at a time
	};

	std::mutex g_ctx_mutex;
	DeviceCtx* g_ctx[kMaxDevices] = {};
	std::atomic<uint64_t> g_launches{0};

	struct DeviceGuard {
		int prev = -1;
		bool active = false;
		int enter(int device) {
			cudaError_t e = cudaGetDevice(&prev);
			if(e != cudaSuccess) return fail_cuda(e, "cudaGetDevice");
			if(prev != device) {
				e = cudaSetDevice(device);
				if(e != cudaSuccess) return fail_cuda(e, "cudaSetDevice");
				active = true;
			}
			return CACAO_OK;
		}
		~DeviceGuard() {
			if(active) cudaSetDevice(prev);
		}
	};

	struct DeviceCtx;
	// A host-memory pipeline that fails half way must not return while copies that read or write the CALLER's
	// buffers are still queued on its streams: the caller is free to release them as soon as it sees the error.
	int drained(DeviceCtx* c, int rc);

	int resolve_device(int device, int* out) {
		int n = 0;
		cudaError_t e = cudaGetDeviceCount(&n);
		if(e != cudaSuccess) return fail_cuda(e, "cudaGetDeviceCount");
		if(n <= 0) return fail(CACAO_E_NO_DEVICE, "no CUDA device is visible; libcacaoimage-b200 has no CPU fallback");
		if(device < 0) {
			e = cudaGetDevice(&device);
			if(e != cudaSuccess) return fail_cuda(e, "cudaGetDevice");
		}
		if(device >= n || device >= kMaxDevices) return fail(CACAO_E_BAD_ARG, "device %d out of range (%d visible)", device, n);
		*out = device;
		return CACAO_OK;
	}

	int drained(DeviceCtx* c, int rc) {
		if(rc == CACAO_OK) return rc; // the success paths end in a synchronise of their own
		const std::string why = t_last_error;
		cudaStreamSynchronize(c->s_in);
		cudaStreamSynchronize(c->s_run);
		cudaStreamSynchronize(c->s_out);
		cudaGetLastError();
		t_last_error = why;
		return rc;
	}

	void destroy_ctx(DeviceCtx* c) {
		int prev = 0;
		cudaGetDevice(&prev);
		cudaSetDevice(c->device);
		cudaDeviceSynchronize();
		for(int s = 0; s < kSlots; ++s) {
			if(c->stage_in[s]) cudaFree(c->stage_in[s]);
			if(c->stage_out[s]) cudaFree(c->stage_out[s]);
			if(c->pin_in[s]) cudaFreeHost(c->pin_in[s]);
			if(c->pin_out[s]) cudaFreeHost(c->pin_out[s]);
			if(c->ev_in[s]) cudaEventDestroy(c->ev_in[s]);
			if(c->ev_run[s]) cudaEventDestroy(c->ev_run[s]);
			if(c->ev_out[s]) cudaEventDestroy(c->ev_out[s]);
		}
		if(c->s_in) cudaStreamDestroy(c->s_in);
		if(c->s_run) cudaStreamDestroy(c->s_run);
		if(c->s_out) cudaStreamDestroy(c->s_out);
		if(c->lut_dev) cudaFree(c->lut_dev);
		if(c->lut2_dev) cudaFree(c->lut2_dev);
		if(c->lanes_dev) cudaFree(c->lanes_dev);
		if(c->checksum_dev) cudaFree(c->checksum_dev);
		cudaSetDevice(prev);
		delete c;
	}

	// Creates everything a device context owns; on failure the caller destroys the partial context.
	int init_ctx(DeviceCtx* c, int dev) {
		c->device = dev;
		cudaDeviceProp prop;
		CU_TRY(cudaGetDeviceProperties(&prop, dev));
		if(prop.major < 10) return fail(CACAO_E_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", dev, prop.major, prop.minor);
		c->sm_count = prop.multiProcessorCount;
		c->smem_optin = prop.sharedMemPerBlockOptin;
		if(!host_lut().two_level_ok) return fail(CACAO_E_BAD_ARG, "internal: two-level 16->8 table does not reproduce the reference table");
		CU_TRY(cudaMalloc(&c->lut_dev, 65536));
		CU_TRY(cudaMemcpy(c->lut_dev, host_lut().table, 65536, cudaMemcpyHostToDevice));
		if(!host_lut().lanes_ok) return fail(CACAO_E_BAD_ARG, "internal: log-indexed 16->8 table does not reproduce the reference table");
		CU_TRY(cudaMalloc(&c->lanes_dev, sizeof(host_lut().lanes)));
		CU_TRY(cudaMemcpy(c->lanes_dev, host_lut().lanes, sizeof(host_lut().lanes), cudaMemcpyHostToDevice));
		CU_TRY(cudaMalloc(&c->lut2_dev, sizeof(host_lut().two_level)));
		CU_TRY(cudaMemcpy(c->lut2_dev, host_lut().two_level, sizeof(host_lut().two_level), cudaMemcpyHostToDevice));
		CU_TRY(cudaMalloc(&c->checksum_dev, sizeof(unsigned long long)));
		CU_TRY(cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking));
		CU_TRY(cudaStreamCreateWithFlags(&c->s_run, cudaStreamNonBlocking));
		CU_TRY(cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking));
		for(int s = 0; s < kSlots; ++s) {
			CU_TRY(cudaEventCreateWithFlags(&c->ev_in[s], cudaEventDisableTiming));
			CU_TRY(cudaEventCreateWithFlags(&c->ev_run[s], cudaEventDisableTiming));
			CU_TRY(cudaEventCreateWithFlags(&c->ev_out[s], cudaEventDisableTiming));
		}
		return CACAO_OK;
	}

	// Returns the context for `device` (creating it on first use) with the device made current.
	int get_ctx(int device, DeviceCtx** out, DeviceGuard& guard) {
		int dev = 0;
		int rc = resolve_device(device, &dev);
		if(rc) return rc;
		rc = guard.enter(dev);
		if(rc) return rc;
		std::lock_guard<std::mutex> lock(g_ctx_mutex);
		if(!g_ctx[dev]) {
			DeviceCtx* c = new DeviceCtx();
			rc = init_ctx(c, dev);
			if(rc) {
				const std::string why = t_last_error; // destroy_ctx makes CUDA calls of its own
				destroy_ctx(c);
				t_last_error = why;
				return rc;
			}
			g_ctx[dev] = c;
		}
		*out = g_ctx[dev];
		return CACAO_OK;
	}

	// Device staging buffers of the host-memory paths.  The chunked conversion pipeline uses all kSlots
	// slots; the single-shot paths (equirect, flip, shift) use slot 0 only, so only it grows for them.
	int ensure_stage(DeviceCtx* c, size_t in_bytes, size_t out_bytes, int slots = kSlots) {
		auto grow = [](uint8_t*& buf, size_t& cap, size_t need, const char* what) -> int {
			if(need <= cap) return CACAO_OK;
			if(buf) cudaFree(buf);
			buf = nullptr;
			cap = 0;
			cudaError_t e = cudaMalloc(&buf, need);
			if(e != cudaSuccess) {
				cudaGetLastError();
				return fail(CACAO_E_NO_MEMORY, "cudaMalloc(%zu) for the %s ring failed: %s", need, what, cudaGetErrorString(e));
			}
			cap = need;
			return CACAO_OK;
		};
		for(int s = 0; s < slots; ++s) {
			int rc = grow(c->stage_in[s], c->stage_in_cap[s], in_bytes, "upload");
			if(rc) return rc;
			rc = grow(c->stage_out[s], c->stage_out_cap[s], out_bytes, "readback");
			if(rc) return rc;
		}
		return CACAO_OK;
	}

	int ensure_pinned(DeviceCtx* c, size_t in_bytes, size_t out_bytes) {
		auto grow = [](uint8_t* (&ring)[kSlots], size_t& cap, size_t need) -> int {
			if(need <= cap) return CACAO_OK;
			for(int s = 0; s < kSlots; ++s) {
				if(ring[s]) cudaFreeHost(ring[s]);
				ring[s] = nullptr;
			}
			cap = 0;
			for(int s = 0; s < kSlots; ++s) {
				cudaError_t e = cudaHostAlloc(reinterpret_cast<void**>(&ring[s]), need, cudaHostAllocPortable);
				if(e != cudaSuccess) return fail(CACAO_E_NO_MEMORY, "cudaHostAlloc(%zu) for the pinned staging ring failed: %s", need, cudaGetErrorString(e));
			}
			cap = need;
			return CACAO_OK;
		};
		int rc = grow(c->pin_in, c->pin_in_cap, in_bytes);
		if(rc) return rc;
		return grow(c->pin_out, c->pin_out_cap, out_bytes);
	}

	// Blocking host->device / device->host copies that stay fast for pageable memory: chunks go through
	// the pinned ring, filled / drained by the copy pool while the previous chunk is on the wire.
	constexpr size_t kXferChunk = 32u << 20;

	int upload_any(DeviceCtx* c, uint8_t* dev_dst, const uint8_t* host_src, size_t bytes, cudaStream_t stream) {
		if(!is_pageable(host_src) || bytes < (1u << 20)) {
			CU_TRY(cudaMemcpyAsync(dev_dst, host_src, bytes, cudaMemcpyHostToDevice, stream));
			CU_TRY(cudaStreamSynchronize(stream));
			return CACAO_OK;
		}
		int rc = ensure_pinned(c, kXferChunk, 0);
		if(rc) return rc;
		size_t done = 0;
		for(size_t k = 0; done < bytes; ++k) {
			const int s = static_cast<int>(k % kSlots);
			const size_t n = std::min(kXferChunk, bytes - done);
			if(k >= kSlots) CU_TRY(cudaEventSynchronize(c->ev_in[s]));
			CopyPool::get().copy(c->pin_in[s], host_src + done, n);
			CU_TRY(cudaMemcpyAsync(dev_dst + done, c->pin_in[s], n, cudaMemcpyHostToDevice, stream));
			CU_TRY(cudaEventRecord(c->ev_in[s], stream));
			done += n;
		}
		CU_TRY(cudaStreamSynchronize(stream));
		return CACAO_OK;
	}

	// One or more device -> host segments, read back as ONE chunked stream: adjacent segments are
	// merged, pinned destinations are DMA'd directly, pageable ones go through the pinned ring with the
	// copy pool draining chunk k-1 while chunk k is on the wire -- across segment borders too, so six
	// separate face buffers cost one pipeline fill, not six.
	struct Segment {
		uint8_t* host;
		const uint8_t* dev;
		size_t bytes;
	};

	int download_segments(DeviceCtx* c, std::vector<Segment> segs, cudaStream_t stream) {
		// merge neighbours that are contiguous on both sides
		std::vector<Segment> merged;
		for(const Segment& sg : segs) {
			if(sg.bytes == 0) continue;
			if(!merged.empty() && merged.back().host + merged.back().bytes == sg.host && merged.back().dev + merged.back().bytes == sg.dev)
				merged.back().bytes += sg.bytes;
			else
				merged.push_back(sg);
		}
		struct Chunk {
			uint8_t* host;
			const uint8_t* dev;
			size_t n;
		};
		std::vector<Chunk> staged; // chunks that travel through the pinned ring
		for(const Segment& sg : merged) {
			if(!is_pageable(sg.host) || sg.bytes < (1u << 20)) {
				CU_TRY(cudaMemcpyAsync(sg.host, sg.dev, sg.bytes, cudaMemcpyDeviceToHost, stream));
				continue;
			}
			for(size_t off = 0; off < sg.bytes; off += kXferChunk) staged.push_back({sg.host + off, sg.dev + off, std::min(kXferChunk, sg.bytes - off)});
		}
		if(!staged.empty()) {
			int rc = ensure_pinned(c, 0, kXferChunk);
			if(rc) return rc;
			for(size_t k = 0; k < staged.size() + 1; ++k) {
				if(k < staged.size()) {
					const int sl = static_cast<int>(k % kSlots);
					CU_TRY(cudaMemcpyAsync(c->pin_out[sl], staged[k].dev, staged[k].n, cudaMemcpyDeviceToHost, stream));
					CU_TRY(cudaEventRecord(c->ev_out[sl], stream));
				}
				if(k >= 1) { // drain the previous chunk while this one is on the wire
					const int sl = static_cast<int>((k - 1) % kSlots);
					CU_TRY(cudaEventSynchronize(c->ev_out[sl]));
					CopyPool::get().copy(staged[k - 1].host, c->pin_out[sl], staged[k - 1].n);
				}
			}
		}
		CU_TRY(cudaStreamSynchronize(stream));
		return CACAO_OK;
	}

	int download_any(DeviceCtx* c, uint8_t* host_dst, const uint8_t* dev_src, size_t bytes, cudaStream_t stream) {
		return download_segments(c, {{host_dst, dev_src, bytes}}, stream);
	}

	// ---- argument checks -----------------------------------------------------------------------
	bool ch_ok(int ch) {
		return ch == 1 || ch == 3 || ch == 4;
	}
	bool fmt_ok(int f) {
		return f >= CACAO_FMT_U8 && f <= CACAO_FMT_F32;
	}

	// true when [a, a+na) and [b, b+nb) share a byte
	bool overlaps(const void* a, uint64_t na, const void* b, uint64_t nb) {
		const uintptr_t x = reinterpret_cast<uintptr_t>(a), y = reinterpret_cast<uintptr_t>(b);
		return na && nb && x < y + nb && y < x + na;
	}

	int check_convert_args(const void* src, void* dst, int src_ch, int dst_ch, int src_fmt, int dst_fmt, int mode, int mem) {
		if(!ch_ok(src_ch) || !ch_ok(dst_ch)) return fail(CACAO_E_BAD_ARG, "channel counts must be 1, 3 or 4 (got %d -> %d)", src_ch, dst_ch);
		if(!fmt_ok(src_fmt) || !fmt_ok(dst_fmt)) return fail(CACAO_E_BAD_ARG, "unknown sample format (%d -> %d)", src_fmt, dst_fmt);
		if(mode != CACAO_MODE_REF_EXACT && mode != CACAO_MODE_FIXED) return fail(CACAO_E_BAD_ARG, "unknown mode %d", mode);
		if(mem != CACAO_MEM_HOST && mem != CACAO_MEM_DEVICE) return fail(CACAO_E_BAD_ARG, "unknown memory kind %d", mem);
		if(src_ch == dst_ch && src_fmt == dst_fmt) return fail(CACAO_E_SAME_LAYOUT, "%s", cacao_img_status_string(CACAO_E_SAME_LAYOUT));
		if(!src || !dst) return fail(CACAO_E_BAD_ARG, "null image pointer");
		return CACAO_OK;
	}

	int enqueue_convert(DeviceCtx* c, const void* src, void* dst, uint64_t pixels, uint64_t ppi, int src_ch, int dst_ch, int src_fmt, int dst_fmt, int mode, cudaStream_t stream, uint32_t flip_w = 0, uint32_t flip_h = 0) {
		ConvertJob job;
		job.flip_w = flip_w;
		job.flip_h = flip_h;
		job.src = src;
		job.dst = dst;
		job.pixels = pixels;
		job.pixels_per_image = ppi ? ppi : 1;
		job.src_ch = src_ch;
		job.dst_ch = dst_ch;
		job.src_fmt = src_fmt;
		job.dst_fmt = dst_fmt;
		job.mode = mode;
		job.lut_dev = c->lut_dev;
		job.lut2_dev = c->lut2_dev;
		job.lanes_dev = c->lanes_dev;
		job.dev.sm_count = c->sm_count;
		job.dev.max_smem_optin = c->smem_optin;
		job.stream = stream;
		job.launches = 0;
		cudaError_t e;
		switch(src_fmt) {
			case CACAO_FMT_U8: e = convert_from_u8(job); break;
			case CACAO_FMT_U16: e = convert_from_u16(job); break;
			case CACAO_FMT_F16: e = convert_from_f16(job); break;
			default: e = convert_from_f32(job); break;
		}
		g_launches.fetch_add(job.launches, std::memory_order_relaxed);
		if(e != cudaSuccess) return fail_cuda(e, "convert kernel launch");
		return CACAO_OK;
	}

	uint64_t zero_copy_limit();

	// Host-memory conversion: chunks flow upload -> kernel -> readback over three streams, so PCIe
	// moves in both directions while the SMs convert the chunk in between.
	// flip_w x flip_h != 0 (cacao_img_convert_flip): the images are flip_w x flip_h pixels and come out bottom
	// row first.  Chunks then are bands of whole rows of one image; a band is converted with its own rows
	// mirrored and lands at the mirrored band of its image, so the pipeline is the same.
	int convert_host_pipeline(DeviceCtx* c, const uint8_t* src, uint8_t* dst, uint64_t pixels, uint64_t ppi, int src_ch, int dst_ch, int src_fmt, int dst_fmt, int mode, uint32_t flip_w = 0, uint32_t flip_h = 0) {
		const uint64_t sbpp = static_cast<uint64_t>(src_ch) * fmt_bytes(src_fmt);
		const uint64_t dbpp = static_cast<uint64_t>(dst_ch) * fmt_bytes(dst_fmt);
		const bool quirk = (src_fmt == CACAO_FMT_U8 || src_fmt == CACAO_FMT_U16) && src_ch == 1 && dst_ch == 4 && mode == CACAO_MODE_REF_EXACT;
		// chunk = a multiple of 8192 pixels (every tile size divides it; chunk byte offsets stay
		// 16-byte aligned) moving about 32 MiB across PCIe in the heavier direction
		static const uint64_t chunk_mib = getenv("CACAO_CHUNK_MIB") ? strtoull(getenv("CACAO_CHUNK_MIB"), nullptr, 10) : 32;
		uint64_t chunk_px = (chunk_mib << 20) / (sbpp > dbpp ? sbpp : dbpp);
		// small jobs still want several chunks in flight so copy-in, DMA, kernel and copy-out overlap
		const uint64_t sixth = (pixels + 5) / 6;
		if(sixth < chunk_px) chunk_px = sixth < 131072 ? 131072 : sixth;
		chunk_px = ((chunk_px + 8191) / 8192) * 8192;
		if(quirk || chunk_px > pixels) chunk_px = pixels; // the pixel-0 quirk needs each image whole
		// chunk k: pixels [src_off, src_off + n) of the source stream go to [dst_off, dst_off + n) of the
		// result; rows != 0: the chunk is `rows` whole rows that the kernel mirrors among themselves
		struct Chunk {
			uint64_t src_off, n, dst_off;
			uint32_t rows;
		};
		const bool flipped = flip_w != 0 && flip_h > 1;
		const bool banded = flipped && chunk_px < pixels; // else one chunk holds every image whole
		const uint64_t band_rows = banded ? std::max<uint64_t>(1, std::min<uint64_t>(flip_h, chunk_px / flip_w)) : 0;
		const uint64_t bands = banded ? (flip_h + band_rows - 1) / band_rows : 0;
		if(banded) chunk_px = band_rows * flip_w;
		const uint64_t chunks = banded ? (pixels / ppi) * bands : (pixels + chunk_px - 1) / chunk_px;
		auto chunk_at = [&](uint64_t k) -> Chunk {
			if(!banded) {
				const uint64_t off = k * chunk_px;
				return {off, (pixels - off < chunk_px) ? pixels - off : chunk_px, off, flipped ? flip_h : 0u};
			}
			const uint64_t img = k / bands, y0 = (k % bands) * band_rows, y1 = std::min<uint64_t>(flip_h, y0 + band_rows);
			return {img * ppi + y0 * flip_w, (y1 - y0) * flip_w, img * ppi + (flip_h - y1) * flip_w, static_cast<uint32_t>(y1 - y0)};
		};
		HostTrace trace(flipped ? "convert_flip" : "convert");
		trace.in_bytes = pixels * sbpp;
		trace.out_bytes = pixels * dbpp;
		std::lock_guard<std::mutex> lock(c->pipe_mutex);
		// Small images (icons, 64x64 textures): two DMA copies and their stream hops cost more than the
		// data.  The kernel reads and writes the pinned staging slots directly over PCIe instead
		// (host memory from cudaHostAlloc is device-addressable under unified addressing): one memcpy
		// in, ONE stream operation, one synchronize, one memcpy out.
		if(pixels * (sbpp + dbpp) <= zero_copy_limit()) {
			trace.path = "zero-copy";
			int rc0 = ensure_pinned(c, std::max<uint64_t>(pixels * sbpp, 64u << 10), std::max<uint64_t>(pixels * dbpp, 64u << 10));
			if(rc0) return rc0;
			std::memcpy(c->pin_in[0], src, pixels * sbpp);
			rc0 = enqueue_convert(c, c->pin_in[0], c->pin_out[0], pixels, ppi, src_ch, dst_ch, src_fmt, dst_fmt, mode, c->s_run, flipped ? flip_w : 0, flipped ? flip_h : 0);
			if(rc0) return rc0;
			CU_TRY(cudaStreamSynchronize(c->s_run));
			std::memcpy(dst, c->pin_out[0], pixels * dbpp);
			return CACAO_OK;
		}
		int rc = ensure_stage(c, chunk_px * sbpp, chunk_px * dbpp);
		if(rc) return rc;
		// pageable buffers (std::vector, numpy) travel through the pinned ring; pinned ones are DMA'd directly
		const bool src_pg = is_pageable(src), dst_pg = is_pageable(dst);
		trace.path = (src_pg || dst_pg) ? "pipeline, pageable" : "pipeline, pinned";
		if(src_pg || dst_pg) {
			rc = ensure_pinned(c, src_pg ? chunk_px * sbpp : 0, dst_pg ? chunk_px * dbpp : 0);
			if(rc) return rc;
		}
		// a single chunk has nothing to overlap with: keep its three steps on one stream and save the
		// cross-stream event hops (tens of microseconds matter for a 64x64 texture)
		cudaStream_t q_in = chunks == 1 ? c->s_run : c->s_in, q_out = chunks == 1 ? c->s_run : c->s_out;
		for(uint64_t k = 0; k < chunks + (kSlots - 1); ++k) {
			if(k < chunks) {
				const int s = static_cast<int>(k % kSlots);
				const Chunk ck = chunk_at(k);
				const uint64_t n = ck.n;
				const uint8_t* up_from = src + ck.src_off * sbpp;
				if(src_pg) {
					if(k >= kSlots) CU_TRY(cudaEventSynchronize(c->ev_in[s])); // the ring slot's last upload has left
					CopyPool::get().copy(c->pin_in[s], src + ck.src_off * sbpp, n * sbpp);
					up_from = c->pin_in[s];
				}
				// the device slot is free once its previous readback has finished
				if(k >= kSlots) CU_TRY(cudaStreamWaitEvent(q_in, c->ev_out[s], 0));
				CU_TRY(cudaMemcpyAsync(c->stage_in[s], up_from, n * sbpp, cudaMemcpyHostToDevice, q_in));
				CU_TRY(cudaEventRecord(c->ev_in[s], q_in));
				if(chunks > 1) CU_TRY(cudaStreamWaitEvent(c->s_run, c->ev_in[s], 0));
				if(k >= kSlots) CU_TRY(cudaStreamWaitEvent(c->s_run, c->ev_out[s], 0));
				rc = enqueue_convert(c, c->stage_in[s], c->stage_out[s], n, quirk ? ppi : n, src_ch, dst_ch, src_fmt, dst_fmt, mode, c->s_run, ck.rows ? flip_w : 0, ck.rows);
				if(rc) return rc;
				CU_TRY(cudaEventRecord(c->ev_run[s], c->s_run));
				if(chunks > 1) CU_TRY(cudaStreamWaitEvent(c->s_out, c->ev_run[s], 0));
				CU_TRY(cudaMemcpyAsync(dst_pg ? c->pin_out[s] : dst + ck.dst_off * dbpp, c->stage_out[s], n * dbpp, cudaMemcpyDeviceToHost, q_out));
				CU_TRY(cudaEventRecord(c->ev_out[s], q_out));
			}
			// drain the chunk issued kSlots-1 iterations ago (its slot is the next one to be reused)
			if(dst_pg && k >= static_cast<uint64_t>(kSlots - 1)) {
				const uint64_t j = k - (kSlots - 1);
				if(j < chunks) {
					const int s = static_cast<int>(j % kSlots);
					const Chunk cj = chunk_at(j);
					CU_TRY(cudaEventSynchronize(c->ev_out[s]));
					CopyPool::get().copy(dst + cj.dst_off * dbpp, c->pin_out[s], cj.n * dbpp);
				}
			}
		}
		CU_TRY(cudaStreamSynchronize(q_out));
		return CACAO_OK;
	}

	// Host-memory resampling of a list of (face, row band) units, pipelined.
	//
	// A unit only reads a window of source rows, so nothing forces "upload everything, resample,
	// read everything back": the source goes up top to bottom in ~32 MiB row chunks (stream s_in), a
	// unit is launched as soon as the last row of its window has been enqueued (s_run waits on the
	// upload event), and its rows start their way back (s_out) while later chunks are still going up.
	// Units are first cut into bands of about 16 MiB of output and ordered by the end of their
	// window, so the +Y face leaves after the first ~30 % of the upload.  PCIe then moves in both
	// directions at once and the call costs about max(upload, readback) instead of their sum.
	// Pinned buffers only; pageable ones and small jobs take the sequential form below.
	// faces[f] = host base of face f (N x N pixels), only needed for faces that units name.
	int equirect_host_pipeline(DeviceCtx* c, const uint8_t* hsrc, uint32_t W, uint32_t H, uint32_t src_row0, uint32_t src_rows, int ch, int fmt, uint8_t* const* faces, uint32_t N, const CubeUnit* units_in, uint32_t n_units_in, int dst_ch, int dst_fmt, uint32_t flags) {
		const uint64_t bpp = static_cast<uint64_t>(ch) * fmt_bytes(fmt);             // source texel
		const uint64_t dbpp = static_cast<uint64_t>(dst_ch) * fmt_bytes(dst_fmt);    // face texel (fused conversion: may differ)
		const uint64_t pitch = static_cast<uint64_t>(W) * bpp;
		const uint64_t face_bytes = static_cast<uint64_t>(N) * N * dbpp;
		struct Piece {
			CubeUnit u;
			uint32_t w0, w1; // source rows [w0, w1) it reads
		};
		std::vector<Piece> pieces;
		const uint64_t row_bytes = static_cast<uint64_t>(N) * dbpp;
		// where rows [r0, r1) of a face are stored: mirrored when the faces come out bottom row first
		const bool flip = (flags & CACAO_EQUIRECT_FLIP_ROWS) != 0;
		auto stored_at = [&](const CubeUnit& u) { return static_cast<uint64_t>(flip ? N - u.row_end : u.row_begin) * row_bytes; };
		const uint32_t band_rows = static_cast<uint32_t>(std::max<uint64_t>(1, std::min<uint64_t>(N, (16ull << 20) / row_bytes)));
		uint32_t lo = UINT32_MAX, hi = 0;
		for(uint32_t k = 0; k < n_units_in; ++k) {
			CubeUnit u = units_in[k];
			if(u.row_end > N) u.row_end = N;
			if(u.face > 5 || u.row_begin >= u.row_end) continue;
			if(!faces[u.face]) return fail(CACAO_E_BAD_ARG, "equirect: no destination for face %u", u.face);
			for(uint32_t r = u.row_begin; r < u.row_end; r += band_rows) {
				Piece p;
				p.u = {u.face, r, std::min(u.row_end, r + band_rows)};
				uint32_t w0 = 0, wn = 0;
				equirect_src_window(W, H, N, 1u << p.u.face, p.u.row_begin, p.u.row_end, &w0, &wn);
				// clip to the rows the caller holds (the kernel clamps the same way)
				// (a window that misses the unit's rows altogether is clamped to its nearest row, like the kernel's taps)
				const uint64_t held_end = static_cast<uint64_t>(src_row0) + src_rows;
				const uint32_t a = static_cast<uint32_t>(std::min<uint64_t>(std::max(w0, src_row0), held_end - 1));
				const uint32_t b = static_cast<uint32_t>(std::min<uint64_t>(static_cast<uint64_t>(w0) + wn, held_end));
				p.w0 = a;
				p.w1 = b > a ? b : a + 1;
				lo = std::min(lo, p.w0);
				hi = std::max(hi, p.w1);
				pieces.push_back(p);
			}
		}
		if(pieces.empty()) return CACAO_OK;
		hi = std::min<uint64_t>(hi, static_cast<uint64_t>(src_row0) + src_rows);
		if(hi <= lo) return fail(CACAO_E_BAD_ARG, "equirect: the source window [%u,+%u) holds none of the rows the units read", src_row0, src_rows);
		std::stable_sort(pieces.begin(), pieces.end(), [](const Piece& a, const Piece& b) { return a.w1 != b.w1 ? a.w1 < b.w1 : a.w0 < b.w0; });

		HostTrace trace("equirect_to_cube");
		trace.in_bytes = static_cast<uint64_t>(hi - lo) * pitch;
		std::lock_guard<std::mutex> lock(c->pipe_mutex);
		int rc = ensure_stage(c, static_cast<uint64_t>(hi - lo) * pitch, 6 * face_bytes, 1);
		if(rc) return rc;
		const uint8_t* up_base = hsrc + static_cast<uint64_t>(lo - src_row0) * pitch;
		const bool src_pg = is_pageable(up_base);
		bool dst_pg = false;
		for(const Piece& p : pieces) dst_pg = dst_pg || is_pageable(faces[p.u.face]);
		// Pageable buffers are staged by host threads in both directions; one orchestrating thread
		// cannot keep both stagings busy at once (measured: 16384x8192 RGBA32F, 89 ms pipelined against
		// 79 ms one after the other), and small jobs only pay for the extra launches and event hops.
		// Those run the three steps in sequence, each still chunk-overlapped inside upload_any /
		// download_any, with whole-face units kept whole for the four-fold kernel.
		uint64_t out_bytes = 0;
		for(const Piece& p : pieces) out_bytes += static_cast<uint64_t>(p.u.row_end - p.u.row_begin) * row_bytes;
		trace.out_bytes = out_bytes;
		uint64_t pipe_min = 32ull << 20, chunk_bytes = kXferChunk;
		if(const char* e = getenv("CACAO_EQ_PIPE_MIN_KIB")) pipe_min = static_cast<uint64_t>(atoll(e)) << 10; // dev switches (dev/time_equirect_host.py)
		if(const char* e = getenv("CACAO_EQ_CHUNK_KIB")) chunk_bytes = std::max<uint64_t>(4096, static_cast<uint64_t>(atoll(e)) << 10);
		const bool small_job = static_cast<uint64_t>(hi - lo) * pitch + out_bytes < pipe_min;
		trace.path = (src_pg || dst_pg) ? "sequential, pageable" : (small_job ? "sequential, small" : "pipelined, pinned");
		if(src_pg || dst_pg || small_job) {
			rc = upload_any(c, c->stage_in[0], up_base, static_cast<uint64_t>(hi - lo) * pitch, c->s_run);
			if(rc) return rc;
			unsigned launches = 0;
			cudaError_t e = equirect_launch_units(c->stage_in[0], c->stage_out[0], ch, fmt, W, H, lo, hi - lo, N, units_in, n_units_in, c->s_run, &launches, dst_ch, dst_fmt, flags);
			g_launches.fetch_add(launches, std::memory_order_relaxed);
			if(e != cudaSuccess) return fail_cuda(e, "equirect kernel launch");
			std::vector<Segment> segs;
			for(uint32_t k = 0; k < n_units_in; ++k) {
				CubeUnit u = units_in[k];
				if(u.row_end > N) u.row_end = N;
				if(u.face > 5 || u.row_begin >= u.row_end) continue;
				const uint64_t off = stored_at(u);
				segs.push_back({faces[u.face] + off, c->stage_out[0] + u.face * face_bytes + off, static_cast<size_t>(u.row_end - u.row_begin) * row_bytes});
			}
			return download_segments(c, segs, c->s_run);
		}
		const uint64_t rows_per_chunk = std::max<uint64_t>(1, chunk_bytes / pitch);
		auto read_back = [&](const Piece& p) -> int {
			const uint64_t off = stored_at(p.u), n = static_cast<uint64_t>(p.u.row_end - p.u.row_begin) * row_bytes;
			CU_TRY(cudaMemcpyAsync(faces[p.u.face] + off, c->stage_out[0] + p.u.face * face_bytes + off, n, cudaMemcpyDeviceToHost, c->s_out));
			return CACAO_OK;
		};

		size_t next = 0;
		uint64_t k = 0;
		for(uint64_t row = lo; row < hi; ++k) {
			const uint64_t rows = std::min<uint64_t>(rows_per_chunk, hi - row);
			const uint8_t* from = hsrc + (row - src_row0) * pitch;
			const int sl = static_cast<int>(k % kSlots);
			CU_TRY(cudaMemcpyAsync(c->stage_in[0] + (row - lo) * pitch, from, rows * pitch, cudaMemcpyHostToDevice, c->s_in));
			CU_TRY(cudaEventRecord(c->ev_in[sl], c->s_in));
			row += rows;
			// every piece whose window is now on its way
			size_t ready_end = next;
			while(ready_end < pieces.size() && pieces[ready_end].w1 <= row) ++ready_end;
			if(row >= hi) ready_end = pieces.size();
			if(ready_end > next) {
				CU_TRY(cudaStreamWaitEvent(c->s_run, c->ev_in[sl], 0));
				std::vector<CubeUnit> batch;
				for(size_t q = next; q < ready_end; ++q) batch.push_back(pieces[q].u);
				unsigned launches = 0;
				cudaError_t e = equirect_launch_units(c->stage_in[0], c->stage_out[0], ch, fmt, W, H, lo, hi - lo, N, batch.data(), static_cast<uint32_t>(batch.size()), c->s_run, &launches, dst_ch, dst_fmt, flags);
				g_launches.fetch_add(launches, std::memory_order_relaxed);
				if(e != cudaSuccess) return fail_cuda(e, "equirect kernel launch");
				CU_TRY(cudaEventRecord(c->ev_run[0], c->s_run));
				CU_TRY(cudaStreamWaitEvent(c->s_out, c->ev_run[0], 0));
				for(size_t q = next; q < ready_end; ++q) {
					rc = read_back(pieces[q]);
					if(rc) return rc;
				}
				next = ready_end;
			}
		}
		CU_TRY(cudaStreamSynchronize(c->s_out));
		return CACAO_OK;
	}

	// One-shot host-memory operation (flip, shift, texel unpack, mip level): small jobs run zero-copy on
	// the pinned staging slots (see convert_host_pipeline), larger ones upload, run and read back through
	// device staging slot 0.  launch(src_dev, dst_dev, stream) enqueues the kernels.
	uint64_t zero_copy_limit() {
		static const uint64_t limit = getenv("CACAO_ZERO_COPY_KIB") ? strtoull(getenv("CACAO_ZERO_COPY_KIB"), nullptr, 10) << 10 : (512ull << 10);
		return limit;
	}

	template<class Launch>
	int run_host_op_body(DeviceCtx* c, const char* what, const void* src, uint64_t in_bytes, void* dst, uint64_t out_bytes, Launch launch);
	template<class Launch>
	int run_host_op(DeviceCtx* c, const char* what, const void* src, uint64_t in_bytes, void* dst, uint64_t out_bytes, Launch launch) {
		return drained(c, run_host_op_body(c, what, src, in_bytes, dst, out_bytes, launch));
	}
	template<class Launch>
	int run_host_op_body(DeviceCtx* c, const char* what, const void* src, uint64_t in_bytes, void* dst, uint64_t out_bytes, Launch launch) {
		HostTrace trace(what);
		trace.in_bytes = in_bytes;
		trace.out_bytes = out_bytes;
		trace.path = in_bytes + out_bytes <= zero_copy_limit() ? "zero-copy" : "sequential";
		std::lock_guard<std::mutex> lock(c->pipe_mutex);
		if(in_bytes + out_bytes <= zero_copy_limit()) {
			int rc = ensure_pinned(c, std::max<uint64_t>(in_bytes, 64u << 10), std::max<uint64_t>(out_bytes, 64u << 10));
			if(rc) return rc;
			std::memcpy(c->pin_in[0], src, in_bytes);
			rc = launch(c->pin_in[0], c->pin_out[0], c->s_run);
			if(rc) return rc;
			CU_TRY(cudaStreamSynchronize(c->s_run));
			std::memcpy(dst, c->pin_out[0], out_bytes);
			return CACAO_OK;
		}
		int rc = ensure_stage(c, in_bytes, out_bytes, 1);
		if(rc) return rc;
		rc = upload_any(c, c->stage_in[0], static_cast<const uint8_t*>(src), in_bytes, c->s_run);
		if(rc) return rc;
		rc = launch(c->stage_in[0], c->stage_out[0], c->s_run);
		if(rc) return rc;
		return download_any(c, static_cast<uint8_t*>(dst), c->stage_out[0], out_bytes, c->s_run);
	}

} // namespace

extern "C" {

int cacao_img_abi_version(void) {
	return CACAO_IMG_ABI_VERSION;
}

const char* cacao_img_last_error(void) {
	return t_last_error.c_str();
}

const char* cacao_img_status_string(int status) {
	switch(status) {
		case CACAO_OK: return "ok";
		// the reference's exception texts, verbatim, so the C++ wrappers throw what callers expect
		case CACAO_E_SAME_LAYOUT: return "Cannot change an image's channel layout to its current layout!";                           // Layout.cpp:9
		case CACAO_E_NOT_16BIT: return "Source image for 16 to 8-bit color conversion does not have a 16-bit color depth!";            // Colordepth.cpp:8
		case CACAO_E_ZERO_DIM: return "Cannot flip an image with zeroed dimensions!";                                                  // Flip.cpp:8
		case CACAO_E_BAD_DEPTH: return "Invalid color depth; only 8 and 16 are allowed.";                                              // Flip.cpp:9
		case CACAO_E_EMPTY: return "Cannot flip an image with a zero-sized data buffer!";                                              // Flip.cpp:10
		case CACAO_E_BAD_ARG: return "invalid argument";
		case CACAO_E_CUDA: return "CUDA error";
		case CACAO_E_NO_DEVICE: return "no usable CUDA device (this library has no CPU fallback)";
		case CACAO_E_NO_MEMORY: return "out of device memory";
	}
	return "unknown status";
}

int cacao_img_device_count(void) {
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if(e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return 0;
	if(e != cudaSuccess) return fail_cuda(e, "cudaGetDeviceCount");
	return n;
}

int cacao_img_init(int device) {
	DeviceGuard g;
	DeviceCtx* c = nullptr;
	return get_ctx(device, &c, g);
}

int cacao_img_shutdown(void) {
	std::lock_guard<std::mutex> lock(g_ctx_mutex);
	for(int d = 0; d < kMaxDevices; ++d) {
		if(g_ctx[d]) {
			destroy_ctx(g_ctx[d]);
			g_ctx[d] = nullptr;
		}
	}
	return CACAO_OK;
}

uint64_t cacao_img_launch_count(void) {
	return g_launches.load(std::memory_order_relaxed);
}

int cacao_img_malloc(int device, void** dev_ptr, uint64_t bytes) {
	if(!dev_ptr) return fail(CACAO_E_BAD_ARG, "null out-pointer");
	DeviceGuard g;
	DeviceCtx* c = nullptr;
	int rc = get_ctx(device, &c, g);
	if(rc) return rc;
	cudaError_t e = cudaMalloc(dev_ptr, bytes ? bytes : 1);
	if(e == cudaErrorMemoryAllocation) {
		cudaGetLastError();
		return fail(CACAO_E_NO_MEMORY, "cudaMalloc(%llu) failed", static_cast<unsigned long long>(bytes));
	}
	if(e != cudaSuccess) return fail_cuda(e, "cudaMalloc");
	return CACAO_OK;
}

int cacao_img_free(int device, void* dev_ptr) {
	DeviceGuard g;
	DeviceCtx* c = nullptr;
	int rc = get_ctx(device, &c, g);
	if(rc) return rc;
	CU_TRY(cudaFree(dev_ptr));
	return CACAO_OK;
}

int cacao_img_host_alloc(void** host_ptr, uint64_t bytes) {
	if(!host_ptr) return fail(CACAO_E_BAD_ARG, "null out-pointer");
	cudaError_t e = cudaHostAlloc(host_ptr, bytes ? bytes : 1, cudaHostAllocPortable);
	if(e != cudaSuccess) return fail_cuda(e, "cudaHostAlloc");
	return CACAO_OK;
}

int cacao_img_host_free(void* host_ptr) {
	CU_TRY(cudaFreeHost(host_ptr));
	return CACAO_OK;
}

int cacao_img_host_register(void* host_ptr, uint64_t bytes) {
	if(!host_ptr || bytes == 0) return fail(CACAO_E_BAD_ARG, "host_register: null pointer or empty range");
	cudaError_t e = cudaHostRegister(host_ptr, bytes, cudaHostRegisterPortable);
	if(e != cudaSuccess) return fail_cuda(e, "cudaHostRegister");
	return CACAO_OK;
}

int cacao_img_host_unregister(void* host_ptr) {
	CU_TRY(cudaHostUnregister(host_ptr));
	return CACAO_OK;
}

int cacao_img_upload(int device, void* dst_dev, const void* src_host, uint64_t bytes, void* stream) {
	DeviceGuard g;
	DeviceCtx* c = nullptr;
	int rc = get_ctx(device, &c, g);
	if(rc) return rc;
	if(bytes >= (1u << 20) && is_pageable(src_host)) {
		// pageable memory cannot be DMA'd asynchronously anyway: stage it through the pinned ring with the
		// host copy pool (3x the driver's own pageable path) and return when it has landed
		std::lock_guard<std::mutex> lock(c->pipe_mutex);
		return upload_any(c, static_cast<uint8_t*>(dst_dev), static_cast<const uint8_t*>(src_host), bytes, static_cast<cudaStream_t>(stream));
	}
	CU_TRY(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream)));
	return CACAO_OK;
}

int cacao_img_download(int device, void* dst_host, const void* src_dev, uint64_t bytes, void* stream) {
	DeviceGuard g;
	DeviceCtx* c = nullptr;
	int rc = get_ctx(device, &c, g);
	if(rc) return rc;
	if(bytes >= (1u << 20) && is_pageable(dst_host)) {
		std::lock_guard<std::mutex> lock(c->pipe_mutex);
		return download_any(c, static_cast<uint8_t*>(dst_host), static_cast<const uint8_t*>(src_dev), bytes, static_cast<cudaStream_t>(stream));
	}
	CU_TRY(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
	return CACAO_OK;
}

int cacao_img_sync(int device, void* stream) {
	DeviceGuard g;
	DeviceCtx* c = nullptr;
	int rc = get_ctx(device, &c, g);
	if(rc) return rc;
	if(stream)
		CU_TRY(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
	else
		CU_TRY(cudaDeviceSynchronize());
	return CACAO_OK;
}

int cacao_img_convert_batch(int device, const void* src, void* dst, uint64_t pixels_per_image, uint64_t count, int src_ch, int dst_ch, int src_fmt, int dst_fmt, int mode, int mem, void* stream) {
	int rc = check_convert_args(src, dst, src_ch, dst_ch, src_fmt, dst_fmt, mode, mem);
	if(rc) return rc;
	DeviceGuard g;
	DeviceCtx* c = nullptr;
	rc = get_ctx(device, &c, g);
	if(rc) return rc;
	const uint64_t pixels = pixels_per_image * count;
	if(pixels == 0) return CACAO_OK;
	// the kernels read through the non-coherent path and promise the compiler disjoint buffers
	if(overlaps(src, pixels * src_ch * fmt_bytes(src_fmt), dst, pixels * dst_ch * fmt_bytes(dst_fmt))) return fail(CACAO_E_BAD_ARG, "convert: source and destination overlap (the conversion is not an in-place operation)");
	if(mem == CACAO_MEM_DEVICE) return enqueue_convert(c, src, dst, pixels, pixels_per_image, src_ch, dst_ch, src_fmt, dst_fmt, mode, static_cast<cudaStream_t>(stream));
	return drained(c, convert_host_pipeline(c, static_cast<const uint8_t*>(src), static_cast<uint8_t*>(dst), pixels, pixels_per_image, src_ch, dst_ch, src_fmt, dst_fmt, mode));
}

int cacao_img_convert_flip(int device, const void* src, void* dst, uint32_t w, uint32_t h, uint64_t count, int src_ch, int dst_ch, int src_fmt, int dst_fmt, int mode, int mem, void* stream) {
	int rc = check_convert_args(src, dst, src_ch, dst_ch, src_fmt, dst_fmt, mode, mem);
	if(rc) return rc;
	if(w == 0 || h == 0) return fail(CACAO_E_ZERO_DIM, "%s", cacao_img_status_string(CACAO_E_ZERO_DIM));
	DeviceGuard g;
	DeviceCtx* c = nullptr;
	rc = get_ctx(device, &c, g);
	if(rc) return rc;
	const uint64_t ppi = static_cast<uint64_t>(w) * h, pixels = ppi * count;
	if(pixels == 0) return CACAO_OK;
	if(overlaps(src, pixels * src_ch * fmt_bytes(src_fmt), dst, pixels * dst_ch * fmt_bytes(dst_fmt))) return fail(CACAO_E_BAD_ARG, "convert_flip: source and destination overlap (a row mirror cannot run in place)");
	if(mem == CACAO_MEM_DEVICE) return enqueue_convert(c, src, dst, pixels, ppi, src_ch, dst_ch, src_fmt, dst_fmt, mode, static_cast<cudaStream_t>(stream), w, h);
	// host memory: the same chunked upload / kernel / readback pipeline, cut into bands of whole rows
	return drained(c, convert_host_pipeline(c, static_cast<const uint8_t*>(src), static_cast<uint8_t*>(dst), pixels, ppi, src_ch, dst_ch, src_fmt, dst_fmt, mode, w, h));
}

int cacao_img_convert(int device, const void* src, void* dst, uint64_t pixels, int src_ch, int dst_ch, int src_fmt, int dst_fmt, int mode, int mem, void* stream) {
	return cacao_img_convert_batch(device, src, dst, pixels, 1, src_ch, dst_ch, src_fmt, dst_fmt, mode, mem, stream);
}

namespace {
	// Mip levels below a level of `count` images of w x h pixels, `ch` 8-bit channels, stored back to back (device
	// memory); the next levels are written behind it.  Stacked images of even height ARE one image of height
	// count * h for a 2x2 box filter (no box straddles two images), so such a level is one launch; odd heights
	// go image by image.
	int enqueue_mips(uint8_t* level, uint32_t w, uint32_t h, uint32_t count, int ch, cudaStream_t stream) {
		while(w > 1 || h > 1) {
			const uint32_t wo = w / 2 ? w / 2 : 1, ho = h / 2 ? h / 2 : 1;
			uint8_t* next = level + static_cast<uint64_t>(count) * w * h * ch;
			unsigned launches = 0, total = 0;
			cudaError_t e = cudaSuccess;
			if(h % 2 == 0 || count == 1) {
				e = downsample2x_launch(level, next, w, count * h, ch, CACAO_FMT_U8, stream, &launches);
				total = launches;
			} else {
				for(uint32_t f = 0; f < count && e == cudaSuccess; ++f) {
					e = downsample2x_launch(level + static_cast<uint64_t>(f) * w * h * ch, next + static_cast<uint64_t>(f) * wo * ho * ch, w, h, ch, CACAO_FMT_U8, stream, &launches);
					total += launches;
				}
			}
			g_launches.fetch_add(total, std::memory_order_relaxed);
			if(e != cudaSuccess) return fail_cuda(e, "mip kernel launch");
			level = next;
			w = wo;
			h = ho;
		}
		return CACAO_OK;
	}

	uint64_t chain_bytes(uint32_t w, uint32_t h, uint32_t count, int ch, bool mips) {
		uint64_t total = static_cast<uint64_t>(count) * w * h * ch;
		if(mips && w && h) {
			while(w > 1 || h > 1) {
				w = w / 2 ? w / 2 : 1;
				h = h / 2 ? h / 2 : 1;
				total += static_cast<uint64_t>(count) * w * h * ch;
			}
		}
		return total;
	}

	// `count` images of w x h pixels stored back to back -> 8-bit samples of dst_ch channels, optionally each image
	// bottom row first; input that already has the destination's form is a copy / a row mirror
	int enqueue_to_u8(DeviceCtx* c, const void* src, void* dst, uint32_t w, uint32_t h, uint64_t count, int ch, int fmt, int dst_ch, bool flip, cudaStream_t stream) {
		const uint64_t ppi = static_cast<uint64_t>(w) * h;
		if(ch == dst_ch && fmt == CACAO_FMT_U8) {
			if(!flip || h == 1) {
				CU_TRY(cudaMemcpyAsync(dst, src, ppi * count * ch, cudaMemcpyDefault, stream));
				return CACAO_OK;
			}
			for(uint64_t k = 0; k < count; ++k) {
				cudaError_t e = flip_rows_launch(static_cast<const uint8_t*>(src) + k * ppi * ch, static_cast<uint8_t*>(dst) + k * ppi * ch, static_cast<uint64_t>(w) * ch, h, c->sm_count, stream);
				if(e != cudaSuccess) return fail_cuda(e, "flip kernel launch");
				g_launches.fetch_add(1, std::memory_order_relaxed);
			}
			return CACAO_OK;
		}
		return enqueue_convert(c, src, dst, ppi * count, ppi, ch, dst_ch, fmt, CACAO_FMT_U8, CACAO_MODE_REF_EXACT, stream, (flip && h > 1) ? w : 0, (flip && h > 1) ? h : 0);
	}

	// One piece of a load chain's level 0: `rows` source rows at `src` (host), converted (and mirrored among
	// themselves when the chain flips) into rows [dst_row, dst_row + rows) of level 0.
	struct ChainPiece {
		const uint8_t* src;
		uint32_t rows, dst_row;
	};

	// Host-memory load chain: pieces flow upload -> convert -> readback over the three streams (piece k+1 goes up
	// while piece k is converted and piece k-1 comes back), the mip levels follow the last piece.  Pageable sources
	// are staged through the pinned ring by the copy pool, pageable destinations drained the same way.
	int run_chain_host(DeviceCtx* c, const char* what, const std::vector<ChainPiece>& pieces, uint32_t w, uint32_t level_rows, int ch, int fmt, int dst_ch, bool flip, bool mips, uint32_t mip_h, uint32_t mip_count, uint8_t* h_out, uint64_t out_total) {
		const uint64_t in_row = static_cast<uint64_t>(w) * ch * fmt_bytes(fmt), out_row = static_cast<uint64_t>(w) * dst_ch;
		uint64_t in_total = 0;
		for(const ChainPiece& p : pieces) in_total += p.rows * in_row;
		HostTrace trace(what);
		trace.in_bytes = in_total;
		trace.out_bytes = out_total;
		std::lock_guard<std::mutex> lock(c->pipe_mutex);
		int rc = ensure_stage(c, in_total, out_total, 1);
		if(rc) return rc;
		bool src_pg = false;
		for(const ChainPiece& p : pieces) src_pg = src_pg || is_pageable(p.src);
		const bool dst_pg = is_pageable(h_out);
		trace.path = (src_pg || dst_pg) ? "piecewise, pageable" : "piecewise, pinned";
		if(src_pg || dst_pg) {
			rc = ensure_pinned(c, src_pg ? kXferChunk : 0, dst_pg ? kXferChunk : 0);
			if(rc) return rc;
		}
		uint8_t* d_in = c->stage_in[0];
		uint8_t* d_out = c->stage_out[0];
		// The first mip levels are made piece by piece too, so that they are computed and read back under the upload
		// of the later pieces instead of after the last one (they are 98 % of the mip bytes): a piece whose first row
		// and row count are multiples of 2^L is, for L halvings, an image of its own to a 2x2 box filter.  `pl` =
		// the levels every piece of this call can make that way (at most kPieceLevels); the rest follow the last piece.
		constexpr uint32_t kPieceLevels = 3;
		struct Level {
			uint64_t off; // byte offset of the level in the output
			uint32_t w;   // its width in pixels
		} lv[kPieceLevels + 2];
		uint32_t pl = 0;
		if(mips) {
			// (a pageable destination is drained after the last piece whatever is done here, and slices under 1 MiB
			// would each be a blocking copy into pageable memory: measured 5.2 -> 12.3 ms; it keeps the old order)
			uint32_t align = 0; // log2 of the largest power of two dividing every piece's rows and first row, and the image height
			for(align = 0; align < kPieceLevels && !dst_pg; ++align) {
				const uint32_t m = (2u << align) - 1u;
				bool ok = (mip_h & m) == 0 && (w >> (align + 1)) >= 1;
				for(const ChainPiece& p : pieces) ok = ok && (p.rows & m) == 0 && (p.dst_row & m) == 0;
				if(!ok) break;
			}
			pl = align;
			lv[0] = {0, w};
			for(uint32_t l = 1; l <= pl + 1; ++l) lv[l] = {lv[l - 1].off + static_cast<uint64_t>(level_rows >> (l - 1)) * lv[l - 1].w * dst_ch, lv[l - 1].w / 2};
		}
		uint64_t ring_use = 0, in_off = 0; // ring slots are reused round robin across pieces
		std::vector<Segment> back;         // pageable destination: drained after the last piece has been issued
		auto read_back = [&](uint64_t off, uint64_t n) -> int {
			if(dst_pg)
				back.push_back({h_out + off, d_out + off, static_cast<size_t>(n)});
			else
				CU_TRY(cudaMemcpyAsync(h_out + off, d_out + off, n, cudaMemcpyDeviceToHost, c->s_out));
			return CACAO_OK;
		};
		for(const ChainPiece& p : pieces) {
			const uint64_t n_in = p.rows * in_row, n_out = p.rows * out_row, out_off = p.dst_row * out_row;
			if(is_pageable(p.src) && n_in >= (1u << 20)) {
				for(uint64_t off = 0; off < n_in; off += kXferChunk, ++ring_use) {
					const int sl = static_cast<int>(ring_use % kSlots);
					const uint64_t n = std::min<uint64_t>(kXferChunk, n_in - off);
					if(ring_use >= kSlots) CU_TRY(cudaEventSynchronize(c->ev_in[sl]));
					CopyPool::get().copy(c->pin_in[sl], p.src + off, n);
					CU_TRY(cudaMemcpyAsync(d_in + in_off + off, c->pin_in[sl], n, cudaMemcpyHostToDevice, c->s_in));
					CU_TRY(cudaEventRecord(c->ev_in[sl], c->s_in));
				}
			} else {
				CU_TRY(cudaMemcpyAsync(d_in + in_off, p.src, n_in, cudaMemcpyHostToDevice, c->s_in));
			}
			CU_TRY(cudaEventRecord(c->ev_run[0], c->s_in)); // "this piece has landed"
			CU_TRY(cudaStreamWaitEvent(c->s_run, c->ev_run[0], 0));
			rc = enqueue_to_u8(c, d_in + in_off, d_out + out_off, w, p.rows, 1, ch, fmt, dst_ch, flip, c->s_run);
			if(rc) return rc;
			for(uint32_t l = 1; l <= pl; ++l) { // the piece's rows of level l from its rows of level l - 1
				const uint64_t src_off = lv[l - 1].off + static_cast<uint64_t>(p.dst_row >> (l - 1)) * lv[l - 1].w * dst_ch;
				const uint64_t dst_off = lv[l].off + static_cast<uint64_t>(p.dst_row >> l) * lv[l].w * dst_ch;
				unsigned launches = 0;
				cudaError_t e = downsample2x_launch(d_out + src_off, d_out + dst_off, lv[l - 1].w, p.rows >> (l - 1), dst_ch, CACAO_FMT_U8, c->s_run, &launches);
				g_launches.fetch_add(launches, std::memory_order_relaxed);
				if(e != cudaSuccess) return fail_cuda(e, "mip kernel launch");
			}
			CU_TRY(cudaEventRecord(c->ev_run[1], c->s_run));
			CU_TRY(cudaStreamWaitEvent(c->s_out, c->ev_run[1], 0));
			rc = read_back(out_off, n_out);
			for(uint32_t l = 1; l <= pl && rc == CACAO_OK; ++l) rc = read_back(lv[l].off + static_cast<uint64_t>(p.dst_row >> l) * lv[l].w * dst_ch, static_cast<uint64_t>(p.rows >> l) * lv[l].w * dst_ch);
			if(rc) return rc;
			in_off += n_in;
		}
		if(mips) {
			// the levels below the last one the pieces made, from that level as a whole
			const uint64_t done = lv[pl + 1].off; // bytes of levels 0 .. pl
			if(done < out_total) {
				rc = enqueue_mips(d_out + lv[pl].off, lv[pl].w, mip_h >> pl, mip_count, dst_ch, c->s_run);
				if(rc) return rc;
				CU_TRY(cudaEventRecord(c->ev_run[2], c->s_run));
				CU_TRY(cudaStreamWaitEvent(c->s_out, c->ev_run[2], 0));
				rc = read_back(done, out_total - done);
				if(rc) return rc;
			}
		}
		if(dst_pg) return download_segments(c, back, c->s_out);
		CU_TRY(cudaStreamSynchronize(c->s_out));
		return CACAO_OK;
	}
}

uint64_t cacao_img_engine_cube_bytes(uint32_t w, uint32_t h, uint32_t flags) {
	return chain_bytes(w, h, 6, 3, (flags & CACAO_CUBE_MIPS) != 0);
}

int cacao_img_engine_cube(int device, const void* const* faces, uint32_t w, uint32_t h, int ch, int fmt, void* dst, uint32_t flags, int mem, void* stream) {
	if(!faces || !dst) return fail(CACAO_E_BAD_ARG, "null pointer");
	for(int f = 0; f < 6; ++f)
		if(!faces[f]) return fail(CACAO_E_BAD_ARG, "engine_cube: face %d is null", f);
	if(!ch_ok(ch)) return fail(CACAO_E_BAD_ARG, "channel count must be 1, 3 or 4 (got %d)", ch);
	if(fmt != CACAO_FMT_U8 && fmt != CACAO_FMT_U16) return fail(CACAO_E_BAD_DEPTH, "%s", cacao_img_status_string(CACAO_E_BAD_DEPTH));
	if(w == 0 || h == 0) return fail(CACAO_E_ZERO_DIM, "%s", cacao_img_status_string(CACAO_E_ZERO_DIM));
	if(flags & ~(CACAO_CUBE_FLIP_ROWS | CACAO_CUBE_MIPS)) return fail(CACAO_E_BAD_ARG, "engine_cube: unknown flags 0x%x", flags);
	if(mem != CACAO_MEM_HOST && mem != CACAO_MEM_DEVICE) return fail(CACAO_E_BAD_ARG, "unknown memory kind %d", mem);
	DeviceGuard g;
	DeviceCtx* c = nullptr;
	int rc = get_ctx(device, &c, g);
	if(rc) return rc;
	const bool flip = (flags & CACAO_CUBE_FLIP_ROWS) != 0, mips = (flags & CACAO_CUBE_MIPS) != 0;
	const uint64_t ppi = static_cast<uint64_t>(w) * h;
	const uint64_t in_face = ppi * ch * fmt_bytes(fmt), out_face = ppi * 3;
	const uint64_t out_total = cacao_img_engine_cube_bytes(w, h, flags);
	for(int f = 0; f < 6; ++f)
		if(overlaps(faces[f], in_face, dst, out_total)) return fail(CACAO_E_BAD_ARG, "engine_cube: face %d overlaps the destination", f);
	if(mem == CACAO_MEM_DEVICE) {
		cudaStream_t s = static_cast<cudaStream_t>(stream);
		bool contiguous = true;
		for(int f = 1; f < 6; ++f) contiguous = contiguous && static_cast<const uint8_t*>(faces[f]) == static_cast<const uint8_t*>(faces[0]) + f * in_face;
		if(contiguous) {
			rc = enqueue_to_u8(c, faces[0], dst, w, h, 6, ch, fmt, 3, flip, s);
			if(rc) return rc;
		} else {
			for(int f = 0; f < 6; ++f) {
				rc = enqueue_to_u8(c, faces[f], static_cast<uint8_t*>(dst) + f * out_face, w, h, 1, ch, fmt, 3, flip, s);
				if(rc) return rc;
			}
		}
		return mips ? enqueue_mips(static_cast<uint8_t*>(dst), w, h, 6, 3, s) : CACAO_OK;
	}
	std::vector<ChainPiece> pieces;
	for(uint32_t f = 0; f < 6; ++f) pieces.push_back({static_cast<const uint8_t*>(faces[f]), h, f * h});
	return drained(c, run_chain_host(c, "engine_cube", pieces, w, 6 * h, ch, fmt, 3, flip, mips, h, 6, static_cast<uint8_t*>(dst), out_total));
}

uint64_t cacao_img_engine_texture_bytes(uint32_t w, uint32_t h, int ch, uint32_t flags) {
	return ch_ok(ch) ? chain_bytes(w, h, 1, ch, (flags & CACAO_CUBE_MIPS) != 0) : 0;
}

int cacao_img_engine_texture(int device, const void* src, uint32_t w, uint32_t h, int ch, int fmt, void* dst, uint32_t flags, int mem, void* stream) {
	if(!src || !dst) return fail(CACAO_E_EMPTY, "%s", cacao_img_status_string(CACAO_E_EMPTY));
	if(!ch_ok(ch)) return fail(CACAO_E_BAD_ARG, "channel count must be 1, 3 or 4 (got %d)", ch);
	if(fmt != CACAO_FMT_U8 && fmt != CACAO_FMT_U16) return fail(CACAO_E_BAD_DEPTH, "%s", cacao_img_status_string(CACAO_E_BAD_DEPTH));
	if(w == 0 || h == 0) return fail(CACAO_E_ZERO_DIM, "%s", cacao_img_status_string(CACAO_E_ZERO_DIM));
	if(flags & ~(CACAO_CUBE_FLIP_ROWS | CACAO_CUBE_MIPS)) return fail(CACAO_E_BAD_ARG, "engine_texture: unknown flags 0x%x", flags);
	if(mem != CACAO_MEM_HOST && mem != CACAO_MEM_DEVICE) return fail(CACAO_E_BAD_ARG, "unknown memory kind %d", mem);
	DeviceGuard g;
	DeviceCtx* c = nullptr;
	int rc = get_ctx(device, &c, g);
	if(rc) return rc;
	const bool flip = (flags & CACAO_CUBE_FLIP_ROWS) != 0, mips = (flags & CACAO_CUBE_MIPS) != 0;
	const uint64_t in_row = static_cast<uint64_t>(w) * ch * fmt_bytes(fmt), out_row = static_cast<uint64_t>(w) * ch;
	const uint64_t out_total = cacao_img_engine_texture_bytes(w, h, ch, flags);
	if(overlaps(src, in_row * h, dst, out_total)) return fail(CACAO_E_BAD_ARG, "engine_texture: source and destination overlap");
	if(mem == CACAO_MEM_DEVICE) {
		cudaStream_t s = static_cast<cudaStream_t>(stream);
		rc = enqueue_to_u8(c, src, dst, w, h, 1, ch, fmt, ch, flip, s);
		if(rc) return rc;
		return mips ? enqueue_mips(static_cast<uint8_t*>(dst), w, h, 1, ch, s) : CACAO_OK;
	}
	// host memory: the image goes through as bands of whole rows (about 32 MiB of the wider side each, at least
	// four so the three streams overlap); a band is mirrored in itself and lands at the mirrored band
	const uint64_t wide = std::max(in_row, out_row);
	uint32_t band = static_cast<uint32_t>(std::max<uint64_t>(1, std::min<uint64_t>(h, kXferChunk / wide)));
	if(h >= 64 && (h + band - 1) / band < 4) band = (h + 3) / 4;
	std::vector<ChainPiece> pieces;
	for(uint32_t y0 = 0; y0 < h; y0 += band) {
		const uint32_t y1 = std::min(h, y0 + band);
		pieces.push_back({static_cast<const uint8_t*>(src) + y0 * in_row, y1 - y0, flip ? h - y1 : y0});
	}
	return drained(c, run_chain_host(c, "engine_texture", pieces, w, h, ch, fmt, ch, flip, mips, h, 1, static_cast<uint8_t*>(dst), out_total));
}

int cacao_img_flip_rows(int device, const void* src, void* dst, uint32_t w, uint32_t h, uint32_t bytes_per_pixel, int mem, void* stream) {
	if(w == 0 || h == 0) return fail(CACAO_E_ZERO_DIM, "%s", cacao_img_status_string(CACAO_E_ZERO_DIM));
	if(bytes_per_pixel == 0) return fail(CACAO_E_BAD_DEPTH, "%s", cacao_img_status_string(CACAO_E_BAD_DEPTH));
	if(!src || !dst) return fail(CACAO_E_EMPTY, "%s", cacao_img_status_string(CACAO_E_EMPTY));
	if(mem != CACAO_MEM_HOST && mem != CACAO_MEM_DEVICE) return fail(CACAO_E_BAD_ARG, "unknown memory kind %d", mem);
	DeviceGuard g;
	DeviceCtx* c = nullptr;
	int rc = get_ctx(device, &c, g);
	if(rc) return rc;
	const uint64_t pitch = static_cast<uint64_t>(w) * bytes_per_pixel;
	if(overlaps(src, pitch * h, dst, pitch * h)) return fail(CACAO_E_BAD_ARG, "flip_rows: source and destination overlap (rows would be read after they were overwritten)");
	if(mem == CACAO_MEM_DEVICE) {
		cudaError_t e = flip_rows_launch(src, dst, pitch, h, c->sm_count, static_cast<cudaStream_t>(stream));
		if(e != cudaSuccess) return fail_cuda(e, "flip kernel launch");
		g_launches.fetch_add(1, std::memory_order_relaxed);
		return CACAO_OK;
	}
	const uint64_t bytes = pitch * h;
	return run_host_op(c, "flip_rows", src, bytes, dst, bytes, [&](uint8_t* s, uint8_t* d, cudaStream_t q) -> int {
		cudaError_t e = flip_rows_launch(s, d, pitch, h, c->sm_count, q);
		if(e != cudaSuccess) return fail_cuda(e, "flip kernel launch");
		g_launches.fetch_add(1, std::memory_order_relaxed);
		return CACAO_OK;
	});
}

int cacao_img_shift_left_u16(int device, const void* src, void* dst, uint64_t samples, uint32_t shift, int mem, void* stream) {
	if(!src || !dst) return fail(CACAO_E_BAD_ARG, "null image pointer");
	if(shift > 15) return fail(CACAO_E_BAD_ARG, "shift %u out of range for 16-bit samples", shift);
	if(mem != CACAO_MEM_HOST && mem != CACAO_MEM_DEVICE) return fail(CACAO_E_BAD_ARG, "unknown memory kind %d", mem);
	DeviceGuard g;
	DeviceCtx* c = nullptr;
	int rc = get_ctx(device, &c, g);
	if(rc) return rc;
	if(samples == 0) return CACAO_OK;
	if(src != dst && overlaps(src, samples * 2, dst, samples * 2)) return fail(CACAO_E_BAD_ARG, "shift_left_u16: buffers may be identical (in place) or disjoint, not partly overlapping");
	const unsigned n_launch = 1u + ((samples % 8) ? 1u : 0u);
	if(mem == CACAO_MEM_DEVICE) {
		cudaError_t e = shift_left_u16_launch(src, dst, samples, shift, static_cast<cudaStream_t>(stream));
		if(e != cudaSuccess) return fail_cuda(e, "shift kernel launch");
		g_launches.fetch_add(n_launch, std::memory_order_relaxed);
		return CACAO_OK;
	}
	return run_host_op(c, "shift_left_u16", src, samples * 2, dst, samples * 2, [&](uint8_t* s, uint8_t* d, cudaStream_t q) -> int {
		cudaError_t e = shift_left_u16_launch(s, d, samples, shift, q);
		if(e != cudaSuccess) return fail_cuda(e, "shift kernel launch");
		g_launches.fetch_add(n_launch, std::memory_order_relaxed);
		return CACAO_OK;
	});
}

int cacao_img_downsample2x(int device, const void* src, void* dst, uint32_t w, uint32_t h, int ch, int fmt, int mem, void* stream) {
	if(!ch_ok(ch) || !fmt_ok(fmt)) return fail(CACAO_E_BAD_ARG, "downsample: bad channel count %d or format %d", ch, fmt);
	if(w == 0 || h == 0) return fail(CACAO_E_ZERO_DIM, "downsample: zero-sized image");
	if(!src || !dst) return fail(CACAO_E_BAD_ARG, "null image pointer");
	if(mem != CACAO_MEM_HOST && mem != CACAO_MEM_DEVICE) return fail(CACAO_E_BAD_ARG, "unknown memory kind %d", mem);
	DeviceGuard g;
	DeviceCtx* c = nullptr;
	int rc = get_ctx(device, &c, g);
	if(rc) return rc;
	const uint64_t bpp = static_cast<uint64_t>(ch) * fmt_bytes(fmt);
	const uint64_t in_bytes = static_cast<uint64_t>(w) * h * bpp;
	const uint64_t out_bytes = static_cast<uint64_t>(w / 2 ? w / 2 : 1) * (h / 2 ? h / 2 : 1) * bpp;
	unsigned launches = 0;
	if(mem == CACAO_MEM_DEVICE) {
		cudaError_t e = downsample2x_launch(src, dst, w, h, ch, fmt, static_cast<cudaStream_t>(stream), &launches);
		g_launches.fetch_add(launches, std::memory_order_relaxed);
		if(e != cudaSuccess) return fail_cuda(e, "downsample kernel launch");
		return CACAO_OK;
	}
	return run_host_op(c, "downsample2x", src, in_bytes, dst, out_bytes, [&](uint8_t* s, uint8_t* d, cudaStream_t q) -> int {
		cudaError_t e = downsample2x_launch(s, d, w, h, ch, fmt, q, &launches);
		g_launches.fetch_add(launches, std::memory_order_relaxed);
		if(e != cudaSuccess) return fail_cuda(e, "downsample kernel launch");
		return CACAO_OK;
	});
}

int cacao_img_unpack_texels(int device, const void* src, void* dst, uint64_t texels, const char* hint, int* out_channels, int mem, void* stream) {
	if(!src || !dst) return fail(CACAO_E_BAD_ARG, "null image pointer");
	if(mem != CACAO_MEM_HOST && mem != CACAO_MEM_DEVICE) return fail(CACAO_E_BAD_ARG, "unknown memory kind %d", mem);
	if(!hint || strnlen(hint, 8) < 8) return fail(CACAO_E_BAD_ARG, "texel format hint must have eight characters (four channel letters, four bit counts)");
	// Model.cpp:250-270: letters -> bit counts, validation, layout
	int src_index[4] = {-1, -1, -1, -1}, bits[4] = {0, 0, 0, 0}, zeroed = 0;
	for(int i = 0; i < 4; ++i) {
		const char* order = "rgba";
		const char* at = strchr(order, hint[i]);
		if(!at || hint[i + 4] < '0' || hint[i + 4] > '8') return fail(CACAO_E_BAD_ARG, "texel format hint '%.8s' is not four of r,g,b,a followed by four digits 0-8", hint);
		const int ch = static_cast<int>(at - order);
		if(src_index[ch] >= 0) return fail(CACAO_E_BAD_ARG, "texel format hint '%.8s' names a channel twice", hint);
		src_index[ch] = i;
		bits[ch] = hint[i + 4] - '0';
	}
	for(int ch = 0; ch < 4; ++ch) zeroed += bits[ch] == 0;
	if(zeroed == 2) return fail(CACAO_E_BAD_ARG, "Invalid zeroed-channel layout for model embedded texture!");                                                         // Model.cpp:259
	if(zeroed == 1 && bits[3] != 0) return fail(CACAO_E_BAD_ARG, "Only the alpha channel may be zero for a model embedded texture layout with only one zeroed channel!"); // :260
	if(zeroed == 4) return fail(CACAO_E_BAD_ARG, "Impossible channel layout for model embedded texture!");                                                               // :269
	TexelPlan plan;
	plan.out_ch = static_cast<uint32_t>(4 - zeroed);
	plan.sel = 0;
	int o = 0;
	for(int ch = 0; ch < 4; ++ch) {
		plan.bits[ch] = 8;
		if(bits[ch] == 0) continue;
		plan.sel |= static_cast<uint32_t>(src_index[ch]) << (4 * o);
		plan.bits[o] = static_cast<uint32_t>(bits[ch]);
		++o;
	}
	for(; o < 4; ++o) plan.sel |= 4u << (4 * o); // byte 0 of the zero operand
	if(out_channels) *out_channels = static_cast<int>(plan.out_ch);
	DeviceGuard g;
	DeviceCtx* c = nullptr;
	int rc = get_ctx(device, &c, g);
	if(rc) return rc;
	if(texels == 0) return CACAO_OK;
	unsigned launches = 0;
	if(mem == CACAO_MEM_DEVICE) {
		cudaError_t e = unpack_texels_launch(src, dst, texels, plan, static_cast<cudaStream_t>(stream), &launches);
		g_launches.fetch_add(launches, std::memory_order_relaxed);
		if(e != cudaSuccess) return fail_cuda(e, "texel unpack kernel launch");
		return CACAO_OK;
	}
	return run_host_op(c, "unpack_texels", src, texels * 4, dst, texels * plan.out_ch, [&](uint8_t* s, uint8_t* d, cudaStream_t q) -> int {
		cudaError_t e = unpack_texels_launch(s, d, texels, plan, q, &launches);
		g_launches.fetch_add(launches, std::memory_order_relaxed);
		if(e != cudaSuccess) return fail_cuda(e, "texel unpack kernel launch");
		return CACAO_OK;
	});
}

int cacao_img_equirect_src_window(uint32_t W, uint32_t H, uint32_t N, uint32_t face_mask, uint32_t row_begin, uint32_t row_end, uint32_t* src_row0, uint32_t* src_rows) {
	if(!src_row0 || !src_rows) return fail(CACAO_E_BAD_ARG, "null out-pointer");
	if(W == 0 || H == 0 || N == 0) return fail(CACAO_E_ZERO_DIM, "equirect: zero-sized image or face");
	equirect_src_window(W, H, N, face_mask & 0x3Fu, row_begin, row_end, src_row0, src_rows);
	return CACAO_OK;
}

int cacao_img_equirect_fold_units(uint32_t N, const cacao_cube_unit* units, uint32_t n_units, cacao_cube_unit* out, uint32_t cap, uint32_t* n_out) {
	if(!n_out || (n_units && !units) || (cap && !out)) return fail(CACAO_E_BAD_ARG, "null pointer");
	if(N == 0) return fail(CACAO_E_ZERO_DIM, "equirect: zero-sized face");
	std::vector<CubeUnit> in;
	for(uint32_t k = 0; k < n_units; ++k) {
		CubeUnit u = {units[k].face, units[k].row_begin, std::min(units[k].row_end, N)};
		if(u.face <= 5 && u.row_begin < u.row_end) in.push_back(u);
	}
	const std::vector<CubeUnit> folded = pair_mirror_units(in, N);
	for(size_t k = 0; k < folded.size() && k < cap; ++k) out[k] = {folded[k].face, folded[k].row_begin, folded[k].row_end};
	*n_out = static_cast<uint32_t>(folded.size());
	return CACAO_OK;
}

int cacao_img_equirect_to_cube(int device, const void* src, uint32_t W, uint32_t H, uint32_t src_row0, uint32_t src_rows, int ch, int fmt, void* dst, uint32_t N, uint32_t face_mask, uint32_t row_begin, uint32_t row_end, int mem, void* stream) {
	if(!ch_ok(ch) || !fmt_ok(fmt)) return fail(CACAO_E_BAD_ARG, "equirect: bad channel count %d or format %d", ch, fmt);
	if(W == 0 || H == 0 || N == 0) return fail(CACAO_E_ZERO_DIM, "equirect: zero-sized image or face");
	if(!src || !dst) return fail(CACAO_E_BAD_ARG, "null image pointer");
	if(mem != CACAO_MEM_HOST && mem != CACAO_MEM_DEVICE) return fail(CACAO_E_BAD_ARG, "unknown memory kind %d", mem);
	if(W > (1u << 24) || H > (1u << 24) || N > (1u << 23)) return fail(CACAO_E_BAD_ARG, "equirect: dimensions beyond exact float range");
	face_mask &= 0x3Fu;
	if(row_end > N) row_end = N;
	if(src_rows == 0 || src_row0 + static_cast<uint64_t>(src_rows) > H) return fail(CACAO_E_BAD_ARG, "equirect: source window [%u,+%u) outside the image (H=%u)", src_row0, src_rows, H);
	if(face_mask == 0 || row_begin >= row_end) return CACAO_OK;
	DeviceGuard g;
	DeviceCtx* c = nullptr;
	int rc = get_ctx(device, &c, g);
	if(rc) return rc;
	const uint64_t bpp = static_cast<uint64_t>(ch) * fmt_bytes(fmt);
	if(mem == CACAO_MEM_DEVICE) {
		if(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15u) != 0) return fail(CACAO_E_BAD_ARG, "equirect: device buffers must be 16-byte aligned");
		cudaError_t e = equirect_launch(src, dst, ch, fmt, W, H, src_row0, src_rows, N, face_mask, row_begin, row_end, static_cast<cudaStream_t>(stream));
		if(e != cudaSuccess) return fail_cuda(e, "equirect kernel launch");
		g_launches.fetch_add(1, std::memory_order_relaxed);
		return CACAO_OK;
	}
	// host memory: a pipelined upload / resample / readback of only the rows involved
	(void)bpp;
	uint8_t* faces[6];
	CubeUnit units[6];
	uint32_t n = 0;
	for(uint32_t f = 0; f < 6; ++f) {
		faces[f] = static_cast<uint8_t*>(dst) + static_cast<uint64_t>(f) * N * N * bpp;
		if(face_mask & (1u << f)) units[n++] = {f, row_begin, row_end};
	}
	return drained(c, equirect_host_pipeline(c, static_cast<const uint8_t*>(src), W, H, src_row0, src_rows, ch, fmt, faces, N, units, n, ch, fmt, 0));
}

namespace {
	int check_equirect_cvt(int ch, int fmt, int dst_ch, int dst_fmt, uint32_t W, uint32_t H, uint32_t N, uint32_t src_row0, uint32_t src_rows, const cacao_cube_unit* units, uint32_t n_units) {
		if(!ch_ok(ch) || !fmt_ok(fmt)) return fail(CACAO_E_BAD_ARG, "equirect: bad channel count %d or format %d", ch, fmt);
		if(!ch_ok(dst_ch) || !fmt_ok(dst_fmt)) return fail(CACAO_E_BAD_ARG, "equirect: bad destination channel count %d or format %d", dst_ch, dst_fmt);
		if(!equirect_cvt_supported(ch, fmt, dst_ch, dst_fmt))
			return fail(CACAO_E_BAD_ARG, "equirect: fused conversion %d ch fmt %d -> %d ch fmt %d is not defined (float sources only; same channels or RGB <-> RGBA): resample, then cacao_img_convert", ch, fmt, dst_ch, dst_fmt);
		if(W == 0 || H == 0 || N == 0) return fail(CACAO_E_ZERO_DIM, "equirect: zero-sized image or face");
		if(W > (1u << 24) || H > (1u << 24) || N > (1u << 23)) return fail(CACAO_E_BAD_ARG, "equirect: dimensions beyond exact float range");
		if(src_rows == 0 || src_row0 + static_cast<uint64_t>(src_rows) > H) return fail(CACAO_E_BAD_ARG, "equirect: source window [%u,+%u) outside the image (H=%u)", src_row0, src_rows, H);
		for(uint32_t u = 0; u < n_units; ++u)
			if(units[u].face > 5) return fail(CACAO_E_BAD_ARG, "equirect: unit %u names face %u", u, units[u].face);
		return CACAO_OK;
	}
	static_assert(sizeof(cacao_cube_unit) == sizeof(CubeUnit), "unit layouts must match");
	// a NULL unit list means the whole cube
	const CubeUnit* units_or_whole_cube(const cacao_cube_unit* units, uint32_t& n_units, uint32_t N, CubeUnit (&whole)[6]) {
		if(units) return reinterpret_cast<const CubeUnit*>(units);
		for(uint32_t f = 0; f < 6; ++f) whole[f] = {f, 0, N};
		n_units = 6;
		return whole;
	}
}

int cacao_img_equirect_to_cube_units_host_cvt(int device, const void* src, uint32_t W, uint32_t H, uint32_t src_row0, uint32_t src_rows, int ch, int fmt, void* const* faces, uint32_t N, int dst_ch, int dst_fmt, uint32_t flags, const cacao_cube_unit* units, uint32_t n_units) {
	if(!src || !faces || (!units && n_units)) return fail(CACAO_E_BAD_ARG, "null pointer");
	if(flags & ~CACAO_EQUIRECT_FLIP_ROWS) return fail(CACAO_E_BAD_ARG, "equirect: unknown flags 0x%x", flags);
	int rc = check_equirect_cvt(ch, fmt, dst_ch, dst_fmt, W, H, N, src_row0, src_rows, units, n_units);
	if(rc) return rc;
	if(units && n_units == 0) return CACAO_OK;
	DeviceGuard g;
	DeviceCtx* c = nullptr;
	rc = get_ctx(device, &c, g);
	if(rc) return rc;
	CubeUnit whole[6];
	const CubeUnit* list = units_or_whole_cube(units, n_units, N, whole);
	return drained(c, equirect_host_pipeline(c, static_cast<const uint8_t*>(src), W, H, src_row0, src_rows, ch, fmt, reinterpret_cast<uint8_t* const*>(faces), N, list, n_units, dst_ch, dst_fmt, flags));
}

int cacao_img_equirect_to_cube_units_host(int device, const void* src, uint32_t W, uint32_t H, uint32_t src_row0, uint32_t src_rows, int ch, int fmt, void* const* faces, uint32_t N, const cacao_cube_unit* units, uint32_t n_units) {
	// here an empty list means "nothing to do" (the _cvt form reads a NULL list as the whole cube)
	static const cacao_cube_unit none = {0, 0, 0};
	return cacao_img_equirect_to_cube_units_host_cvt(device, src, W, H, src_row0, src_rows, ch, fmt, faces, N, ch, fmt, 0, (units || n_units) ? units : &none, n_units);
}

int cacao_img_equirect_to_cube_units_cvt(int device, const void* src, uint32_t W, uint32_t H, uint32_t src_row0, uint32_t src_rows, int ch, int fmt, void* dst, uint32_t N, int dst_ch, int dst_fmt, uint32_t flags, const cacao_cube_unit* units, uint32_t n_units, void* stream) {
	if(!src || !dst || (!units && n_units)) return fail(CACAO_E_BAD_ARG, "null pointer");
	if(flags & ~CACAO_EQUIRECT_FLIP_ROWS) return fail(CACAO_E_BAD_ARG, "equirect: unknown flags 0x%x", flags);
	int rc = check_equirect_cvt(ch, fmt, dst_ch, dst_fmt, W, H, N, src_row0, src_rows, units, n_units);
	if(rc) return rc;
	if(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15u) != 0) return fail(CACAO_E_BAD_ARG, "equirect: device buffers must be 16-byte aligned");
	if(units && n_units == 0) return CACAO_OK;
	DeviceGuard g;
	DeviceCtx* c = nullptr;
	rc = get_ctx(device, &c, g);
	if(rc) return rc;
	CubeUnit whole[6];
	const CubeUnit* list = units_or_whole_cube(units, n_units, N, whole);
	unsigned launches = 0;
	cudaError_t e = equirect_launch_units(src, dst, ch, fmt, W, H, src_row0, src_rows, N, list, n_units, static_cast<cudaStream_t>(stream), &launches, dst_ch, dst_fmt, flags);
	g_launches.fetch_add(launches, std::memory_order_relaxed);
	if(e != cudaSuccess) return fail_cuda(e, "equirect kernel launch");
	return CACAO_OK;
}

int cacao_img_equirect_to_cube_units(int device, const void* src, uint32_t W, uint32_t H, uint32_t src_row0, uint32_t src_rows, int ch, int fmt, void* dst, uint32_t N, const cacao_cube_unit* units, uint32_t n_units, void* stream) {
	static const cacao_cube_unit none = {0, 0, 0};
	return cacao_img_equirect_to_cube_units_cvt(device, src, W, H, src_row0, src_rows, ch, fmt, dst, N, ch, fmt, 0, (units || n_units) ? units : &none, n_units, stream);
}

int cacao_img_synth(int device, void* dst_dev, uint64_t first_sample, uint64_t samples, int fmt, int ch, uint32_t seed, void* stream) {
	if(!fmt_ok(fmt) || !ch_ok(ch) || !dst_dev) return fail(CACAO_E_BAD_ARG, "synth: bad argument");
	DeviceGuard g;
	DeviceCtx* c = nullptr;
	int rc = get_ctx(device, &c, g);
	if(rc) return rc;
	if(samples == 0) return CACAO_OK;
	cudaError_t e = synth_launch(dst_dev, first_sample, samples, fmt, ch, seed, c->sm_count, static_cast<cudaStream_t>(stream));
	if(e != cudaSuccess) return fail_cuda(e, "synth kernel launch");
	g_launches.fetch_add(1, std::memory_order_relaxed);
	return CACAO_OK;
}

int cacao_img_checksum(int device, const void* src_dev, uint64_t bytes, uint64_t* out, void* stream) {
	if(!src_dev || !out || (bytes & 3u)) return fail(CACAO_E_BAD_ARG, "checksum: null pointer or size not a multiple of 4");
	DeviceGuard g;
	DeviceCtx* c = nullptr;
	int rc = get_ctx(device, &c, g);
	if(rc) return rc;
	std::lock_guard<std::mutex> lock(c->pipe_mutex);
	cudaStream_t s = static_cast<cudaStream_t>(stream);
	CU_TRY(cudaMemsetAsync(c->checksum_dev, 0, sizeof(unsigned long long), s));
	if(bytes) {
		cudaError_t e = checksum_launch(src_dev, bytes / 4, c->checksum_dev, c->sm_count, s);
		if(e != cudaSuccess) return fail_cuda(e, "checksum kernel launch");
		g_launches.fetch_add(1, std::memory_order_relaxed);
	}
	unsigned long long host = 0;
	CU_TRY(cudaMemcpyAsync(&host, c->checksum_dev, sizeof host, cudaMemcpyDeviceToHost, s));
	CU_TRY(cudaStreamSynchronize(s));
	*out = host;
	return CACAO_OK;
}

// ---- device memory that another API or process takes over without a copy -------------------------
namespace {
	// The CUDA driver's virtual-memory calls, fetched at run time like cuTensorMapEncodeTiled in equirect_tma.cu
	// (the library links the runtime statically and has no link-time dependency on libcuda).
	struct VmmApi {
		CUresult (*get_granularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
		CUresult (*create)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
		CUresult (*release)(CUmemGenericAllocationHandle) = nullptr;
		CUresult (*reserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
		CUresult (*address_free)(CUdeviceptr, size_t) = nullptr;
		CUresult (*map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
		CUresult (*unmap)(CUdeviceptr, size_t) = nullptr;
		CUresult (*set_access)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
		CUresult (*export_handle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
		CUresult (*import_handle)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType) = nullptr;
		bool ok = false;
		VmmApi() {
			auto get = [](const char* name, auto& fn) {
				void* p = nullptr;
				cudaDriverEntryPointQueryResult q;
				if(cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
				fn = reinterpret_cast<std::remove_reference_t<decltype(fn)>>(p);
				return p != nullptr;
			};
			ok = get("cuMemGetAllocationGranularity", get_granularity) & get("cuMemCreate", create) & get("cuMemRelease", release) & get("cuMemAddressReserve", reserve) &
				 get("cuMemAddressFree", address_free) & get("cuMemMap", map) & get("cuMemUnmap", unmap) & get("cuMemSetAccess", set_access) &
				 get("cuMemExportToShareableHandle", export_handle) & get("cuMemImportFromShareableHandle", import_handle);
		}
	};
	const VmmApi& vmm() {
		static const VmmApi api;
		return api;
	}
	struct Shareable {
		CUmemGenericAllocationHandle handle;
		size_t bytes;
	};
	std::mutex g_shareable_mutex;
	std::map<void*, Shareable> g_shareable;

	CUmemAllocationProp shareable_prop(int device) {
		CUmemAllocationProp prop = {};
		prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
		prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
		prop.location.id = device;
		prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
		return prop;
	}

	// reserve + map + allow read/write from `device`; on failure everything taken so far is given back
	int map_shareable(int device, CUmemGenericAllocationHandle handle, size_t bytes, void** dev_ptr) {
		const VmmApi& v = vmm();
		CUdeviceptr va = 0;
		if(v.reserve(&va, bytes, 0, 0, 0) != CUDA_SUCCESS) {
			v.release(handle);
			return fail(CACAO_E_NO_MEMORY, "cuMemAddressReserve(%llu) failed", static_cast<unsigned long long>(bytes));
		}
		CUmemAccessDesc access = {};
		access.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
		access.location.id = device;
		access.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
		if(v.map(va, bytes, 0, handle, 0) != CUDA_SUCCESS) {
			v.address_free(va, bytes);
			v.release(handle);
			return fail(CACAO_E_CUDA, "cuMemMap failed");
		}
		if(v.set_access(va, bytes, &access, 1) != CUDA_SUCCESS) {
			v.unmap(va, bytes);
			v.address_free(va, bytes);
			v.release(handle);
			return fail(CACAO_E_CUDA, "cuMemSetAccess failed");
		}
		*dev_ptr = reinterpret_cast<void*>(va);
		std::lock_guard<std::mutex> lock(g_shareable_mutex);
		g_shareable[*dev_ptr] = {handle, bytes};
		return CACAO_OK;
	}
}

int cacao_img_malloc_shareable(int device, void** dev_ptr, uint64_t bytes, int* fd, uint64_t* mapped_bytes) {
	if(!dev_ptr || !fd || !mapped_bytes) return fail(CACAO_E_BAD_ARG, "null out-pointer");
	DeviceGuard g;
	DeviceCtx* c = nullptr;
	int rc = get_ctx(device, &c, g);
	if(rc) return rc;
	const VmmApi& v = vmm();
	if(!v.ok) return fail(CACAO_E_CUDA, "the CUDA driver lacks the virtual-memory API");
	const CUmemAllocationProp prop = shareable_prop(c->device);
	size_t gran = 0;
	if(v.get_granularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM) != CUDA_SUCCESS || gran == 0) return fail(CACAO_E_CUDA, "cuMemGetAllocationGranularity failed");
	const size_t size = ((bytes ? bytes : 1) + gran - 1) / gran * gran;
	CUmemGenericAllocationHandle handle = 0;
	if(v.create(&handle, size, &prop, 0) != CUDA_SUCCESS) return fail(CACAO_E_NO_MEMORY, "cuMemCreate(%llu) failed", static_cast<unsigned long long>(size));
	int out_fd = -1;
	if(v.export_handle(&out_fd, handle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0) != CUDA_SUCCESS) {
		v.release(handle);
		return fail(CACAO_E_CUDA, "cuMemExportToShareableHandle failed");
	}
	rc = map_shareable(c->device, handle, size, dev_ptr);
	if(rc) {
		close(out_fd);
		return rc;
	}
	*fd = out_fd;
	*mapped_bytes = size;
	return CACAO_OK;
}

int cacao_img_import_shareable(int device, int fd, uint64_t mapped_bytes, void** dev_ptr) {
	if(!dev_ptr || fd < 0 || mapped_bytes == 0) return fail(CACAO_E_BAD_ARG, "import_shareable: null out-pointer, bad descriptor or empty size");
	DeviceGuard g;
	DeviceCtx* c = nullptr;
	int rc = get_ctx(device, &c, g);
	if(rc) return rc;
	const VmmApi& v = vmm();
	if(!v.ok) return fail(CACAO_E_CUDA, "the CUDA driver lacks the virtual-memory API");
	CUmemGenericAllocationHandle handle = 0;
	if(v.import_handle(&handle, reinterpret_cast<void*>(static_cast<uintptr_t>(fd)), CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR) != CUDA_SUCCESS) return fail(CACAO_E_CUDA, "cuMemImportFromShareableHandle failed (descriptor %d)", fd);
	return map_shareable(c->device, handle, static_cast<size_t>(mapped_bytes), dev_ptr);
}

int cacao_img_free_shareable(int device, void* dev_ptr) {
	DeviceGuard g;
	DeviceCtx* c = nullptr;
	int rc = get_ctx(device, &c, g);
	if(rc) return rc;
	Shareable sh;
	{
		std::lock_guard<std::mutex> lock(g_shareable_mutex);
		auto it = g_shareable.find(dev_ptr);
		if(it == g_shareable.end()) return fail(CACAO_E_BAD_ARG, "free_shareable: not a pointer from malloc_shareable / import_shareable");
		sh = it->second;
		g_shareable.erase(it);
	}
	CU_TRY(cudaDeviceSynchronize());
	const VmmApi& v = vmm();
	const CUdeviceptr va = reinterpret_cast<CUdeviceptr>(dev_ptr);
	if(v.unmap(va, sh.bytes) != CUDA_SUCCESS || v.address_free(va, sh.bytes) != CUDA_SUCCESS || v.release(sh.handle) != CUDA_SUCCESS) return fail(CACAO_E_CUDA, "unmapping a shareable allocation failed");
	return CACAO_OK;
}

// ---- one buffer, several processes: the device-side gather of a sharded job ---------------------------
// One process per GPU is how a box is driven (bench.py under torchrun, ce-cubetool --gpus in threads).  A rank that
// opens another rank's buffer gets a pointer to the SAME memory; handed to any CACAO_MEM_DEVICE call as `dst`, the
// kernel's own stores then cross NVLink and land in place -- a sharded cubemap is assembled on one GPU by the
// resampling launches themselves, with no gather step and no intermediate buffer.
int cacao_img_ipc_export(int device, void* dev_ptr, unsigned char* handle64, uint64_t* offset) {
	if(!dev_ptr || !handle64 || !offset) return fail(CACAO_E_BAD_ARG, "ipc_export: null pointer");
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the handle is 64 opaque bytes");
	DeviceGuard g;
	DeviceCtx* c = nullptr;
	int rc = get_ctx(device, &c, g);
	if(rc) return rc;
	// the handle names the whole allocation: find where in it the pointer lies
	using GetRange = CUresult (*)(CUdeviceptr*, size_t*, CUdeviceptr);
	static const GetRange get_range = [] {
		void* p = nullptr;
		cudaDriverEntryPointQueryResult q;
		if(cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
		return reinterpret_cast<GetRange>(p);
	}();
	CUdeviceptr base = 0;
	size_t size = 0;
	if(!get_range || get_range(&base, &size, reinterpret_cast<CUdeviceptr>(dev_ptr)) != CUDA_SUCCESS) return fail(CACAO_E_BAD_ARG, "ipc_export: not a device allocation of this process");
	cudaIpcMemHandle_t h;
	CU_TRY(cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(base)));
	std::memcpy(handle64, &h, 64);
	*offset = reinterpret_cast<CUdeviceptr>(dev_ptr) - base;
	return CACAO_OK;
}

int cacao_img_ipc_import(int device, const unsigned char* handle64, uint64_t offset, void** dev_ptr) {
	if(!handle64 || !dev_ptr) return fail(CACAO_E_BAD_ARG, "ipc_import: null pointer");
	DeviceGuard g;
	DeviceCtx* c = nullptr;
	int rc = get_ctx(device, &c, g);
	if(rc) return rc;
	cudaIpcMemHandle_t h;
	std::memcpy(&h, handle64, 64);
	void* base = nullptr;
	CU_TRY(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess)); // peer access to the owner's GPU comes with it
	*dev_ptr = static_cast<uint8_t*>(base) + offset;
	return CACAO_OK;
}

int cacao_img_ipc_close(int device, void* dev_ptr, uint64_t offset) {
	if(!dev_ptr) return fail(CACAO_E_BAD_ARG, "ipc_close: null pointer");
	DeviceGuard g;
	DeviceCtx* c = nullptr;
	int rc = get_ctx(device, &c, g);
	if(rc) return rc;
	CU_TRY(cudaDeviceSynchronize());
	CU_TRY(cudaIpcCloseMemHandle(static_cast<uint8_t*>(dev_ptr) - offset));
	return CACAO_OK;
}

// Same idea inside ONE process that drives several devices (EquirectToCubemap(..., devices)): let kernels on
// `device` read and write memory of `peer`.  Already enabled / same device: success.
int cacao_img_enable_peer_access(int device, int peer) {
	DeviceGuard g;
	DeviceCtx* c = nullptr;
	int rc = get_ctx(device, &c, g);
	if(rc) return rc;
	DeviceGuard g2;
	DeviceCtx* cp = nullptr;
	rc = get_ctx(peer, &cp, g2);
	if(rc) return rc;
	if(c->device == cp->device) return CACAO_OK;
	int can = 0;
	CU_TRY(cudaSetDevice(c->device));
	CU_TRY(cudaDeviceCanAccessPeer(&can, c->device, cp->device));
	if(!can) return fail(CACAO_E_CUDA, "device %d cannot access device %d", c->device, cp->device);
	cudaError_t e = cudaDeviceEnablePeerAccess(cp->device, 0);
	if(e == cudaErrorPeerAccessAlreadyEnabled) {
		cudaGetLastError();
		return CACAO_OK;
	}
	if(e != cudaSuccess) return fail_cuda(e, "cudaDeviceEnablePeerAccess");
	return CACAO_OK;
}

int cacao_img_get_lut(uint8_t* out65536) {
	if(!out65536) return fail(CACAO_E_BAD_ARG, "null out-pointer");
	std::memcpy(out65536, host_lut().table, 65536);
	return CACAO_OK;
}

} // extern "C"
