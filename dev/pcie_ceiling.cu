// dev: the host<->device copy ceiling of this box -- plain pinned cudaMemcpyAsync, no kernels, no library.
//
//   nvcc -O2 -o dev/pcie_ceiling dev/pcie_ceiling.cu && dev/pcie_ceiling [max_gpus]
//
// For g = 1, 2, 4, 8 concurrent GPUs (one PROCESS per GPU, forked, started together through a barrier in
// shared memory) each process moves bench.py's e2e step -- 2.12 GB up, 4.25 GB down, pinned host buffers of
// its own -- three ways: upload alone, readback alone, both at once on two streams.  The "both" line is the
// roofline of cacao_img_convert_batch(CACAO_MEM_HOST): nothing that goes through host memory can beat it.
#include <cuda_runtime.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

struct Shared {
	std::atomic<int> arrived[64];
	double up[8], down[8], both[8];
};

static void barrier(Shared* sh, int slot, int n) {
	sh->arrived[slot].fetch_add(1);
	while(sh->arrived[slot].load() < n) usleep(50);
}

#define CK(x)                                                                          \
	do {                                                                               \
		cudaError_t e = (x);                                                           \
		if(e != cudaSuccess) {                                                         \
			fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e));                    \
			_exit(2);                                                                  \
		}                                                                              \
	} while(0)

static double now() {
	return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static void child(Shared* sh, int g, int n, int slot0) {
	const size_t in_bytes = 1920ull * 1080 * 256 * 4, out_bytes = in_bytes * 2;
	CK(cudaSetDevice(g));
	void *hin, *hout, *din, *dout;
	CK(cudaHostAlloc(&hin, in_bytes, cudaHostAllocPortable));
	CK(cudaHostAlloc(&hout, out_bytes, cudaHostAllocPortable));
	memset(hin, 1, in_bytes);
	memset(hout, 2, out_bytes);
	CK(cudaMalloc(&din, in_bytes));
	CK(cudaMalloc(&dout, out_bytes));
	CK(cudaMemset(dout, 3, out_bytes));
	cudaStream_t s0, s1;
	CK(cudaStreamCreateWithFlags(&s0, cudaStreamNonBlocking));
	CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
	CK(cudaMemcpyAsync(din, hin, in_bytes, cudaMemcpyHostToDevice, s0));
	CK(cudaMemcpyAsync(hout, dout, out_bytes, cudaMemcpyDeviceToHost, s1));
	CK(cudaDeviceSynchronize());
	const int reps = 3;
	double t;
	barrier(sh, slot0 + 0, n);
	t = now();
	for(int r = 0; r < reps; ++r) CK(cudaMemcpyAsync(din, hin, in_bytes, cudaMemcpyHostToDevice, s0));
	CK(cudaDeviceSynchronize());
	sh->up[g] = reps * in_bytes / (now() - t) / 1e9;
	barrier(sh, slot0 + 1, n);
	t = now();
	for(int r = 0; r < reps; ++r) CK(cudaMemcpyAsync(hout, dout, out_bytes, cudaMemcpyDeviceToHost, s1));
	CK(cudaDeviceSynchronize());
	sh->down[g] = reps * out_bytes / (now() - t) / 1e9;
	barrier(sh, slot0 + 2, n);
	t = now();
	for(int r = 0; r < reps; ++r) {
		CK(cudaMemcpyAsync(din, hin, in_bytes, cudaMemcpyHostToDevice, s0));
		CK(cudaMemcpyAsync(hout, dout, out_bytes, cudaMemcpyDeviceToHost, s1));
	}
	CK(cudaDeviceSynchronize());
	sh->both[g] = reps * (in_bytes + out_bytes) / (now() - t) / 1e9; // seconds for one e2e step = (in+out)/this
	barrier(sh, slot0 + 3, n);
	_exit(0);
}

int main(int argc, char** argv) {
	int max_gpus = argc > 1 ? atoi(argv[1]) : 8;
	Shared* sh = static_cast<Shared*>(mmap(nullptr, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0));
	memset(sh, 0, sizeof(Shared));
	// device count in a child: the parent must not create a CUDA context before it forks
	int pipefd[2];
	if(pipe(pipefd)) return 1;
	if(fork() == 0) {
		int n = 0;
		cudaGetDeviceCount(&n);
		if(write(pipefd[1], &n, sizeof n) != sizeof n) _exit(1);
		_exit(0);
	}
	int have = 0;
	if(read(pipefd[0], &have, sizeof have) != sizeof have) return 1;
	wait(nullptr);
	printf("pinned cudaMemcpyAsync ceiling, %d GPU(s) visible; per process: 2.123 GB up, 4.247 GB down (bench.py's e2e step)\n", have);
	printf("%5s  %22s  %22s  %34s  %s\n", "gpus", "upload alone GB/s", "readback alone GB/s", "both at once GB/s (up+down bytes)", "=> e2e step ms, Gpixel/s aggregate");
	int slot = 0;
	for(int n = 1; n <= max_gpus && n <= have; n *= 2) {
		for(int g = 0; g < n; ++g)
			if(fork() == 0) child(sh, g, n, slot);
		for(int g = 0; g < n; ++g) wait(nullptr);
		slot += 4;
		double up = 0, down = 0, both = 0, worst = 1e30;
		for(int g = 0; g < n; ++g) {
			up += sh->up[g];
			down += sh->down[g];
			both += sh->both[g];
			worst = sh->both[g] < worst ? sh->both[g] : worst;
		}
		const double step_ms = (1920.0 * 1080 * 256 * 12) / worst / 1e6; // slowest rank bounds the step
		printf("%5d  %10.1f (%5.1f/gpu)  %10.1f (%5.1f/gpu)  %16.1f (%5.1f/gpu, min %5.1f)  %.1f ms, %.2f Gpixel/s\n", n, up, up / n, down, down / n, both, both / n, worst, step_ms, n * 1920.0 * 1080 * 256 / step_ms / 1e6);
		fflush(stdout);
	}
	return 0;
}
