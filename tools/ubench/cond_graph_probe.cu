// Probe: CUDA-graph conditional nodes built during stream capture (WHILE body captured on a second stream, cluster kernel
// inside the body, nested IF -> WHILE), launched from several streams at once.  Prints what works on this driver.
#include <cuda_runtime.h>
#include <stdio.h>
#include <vector>
#include <chrono>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("FAIL %s line %d: %s\n", #x, __LINE__, cudaGetErrorString(e_)); return 1; } } while (0)

__global__ void step_kernel(int* counter, int limit, cudaGraphConditionalHandle h, int use) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const int c = ++*counter;
    if (use) cudaGraphSetConditional(h, c < limit ? 1u : 0u);
  }
}
__global__ void set_kernel(const int* flag, cudaGraphConditionalHandle h) {
  if (threadIdx.x == 0) cudaGraphSetConditional(h, *flag ? 1u : 0u);
}
__global__ void __cluster_dims__(4, 1, 1) cluster_kernel(int* acc) {
  if (threadIdx.x == 0) atomicAdd(acc, 1);
}
__global__ void work_kernel(double* x, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = x[i] * 1.0000001 + 1e-9;
}
__global__ void empty_kernel(const int* flag) { if (flag && *flag == 12345) printf("x"); }

static int add_cond(cudaStream_t s, cudaGraphConditionalHandle h, cudaGraphConditionalNodeType type, cudaGraph_t* body) {
  cudaStreamCaptureStatus st; unsigned long long id; cudaGraph_t g; const cudaGraphNode_t* deps; size_t nd;
  CK(cudaStreamGetCaptureInfo_v2(s, &st, &id, &g, &deps, &nd));
  cudaGraphNodeParams p = {};
  p.type = cudaGraphNodeTypeConditional;
  p.conditional.handle = h; p.conditional.type = type; p.conditional.size = 1;
  cudaGraphNode_t node;
  CK(cudaGraphAddNode(&node, g, deps, nd, &p));
  *body = p.conditional.phGraph_out[0];
  CK(cudaStreamUpdateCaptureDependencies(s, &node, 1, cudaStreamSetCaptureDependencies));
  return 0;
}

int main() {
  int* d; double* x;
  CK(cudaMalloc(&d, 64 * sizeof(int))); CK(cudaMemset(d, 0, 64 * sizeof(int)));
  CK(cudaMalloc(&x, 1 << 20));
  cudaStream_t s, b, b2;
  CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&b, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&b2, cudaStreamNonBlocking));
  // ---- test 1: WHILE via capture, cluster kernel in body
  {
    CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    cudaStreamCaptureStatus st; unsigned long long id; cudaGraph_t g; const cudaGraphNode_t* deps; size_t nd;
    CK(cudaStreamGetCaptureInfo_v2(s, &st, &id, &g, &deps, &nd));
    cudaGraphConditionalHandle h;
    CK(cudaGraphConditionalHandleCreate(&h, g, 0, cudaGraphCondAssignDefault));
    step_kernel<<<1, 32, 0, s>>>(d, 5, h, 1);            // counter 1, cond = 1
    cudaGraph_t body;
    if (add_cond(s, h, cudaGraphCondTypeWhile, &body)) return 1;
    CK(cudaStreamBeginCaptureToGraph(b, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
    work_kernel<<<64, 256, 0, b>>>(x, 1 << 14);
    cluster_kernel<<<8, 32, 0, b>>>(d + 1);
    step_kernel<<<1, 32, 0, b>>>(d, 5, h, 1);
    cudaGraph_t dummy;
    CK(cudaStreamEndCapture(b, &dummy));
    step_kernel<<<1, 32, 0, s>>>(d + 2, 0, h, 0);        // after the loop
    cudaGraph_t graph;
    CK(cudaStreamEndCapture(s, &graph));
    cudaGraphExec_t ex;
    CK(cudaGraphInstantiate(&ex, graph, 0));
    for (int r = 0; r < 3; ++r) { CK(cudaMemsetAsync(d, 0, 16, s)); CK(cudaGraphLaunch(ex, s)); }
    CK(cudaStreamSynchronize(s));
    int hst[4]; CK(cudaMemcpy(hst, d, 16, cudaMemcpyDeviceToHost));
    printf("test1 WHILE+cluster: counter %d (want 5) cluster hits %d (want 3*4*8=96) after %d (want 3)\n", hst[0], hst[1], hst[2]);
  }
  // ---- test 2: nested IF -> WHILE
  {
    CK(cudaMemset(d, 0, 64 * sizeof(int)));
    int one = 1; CK(cudaMemcpy(d + 8, &one, 4, cudaMemcpyHostToDevice));
    CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    cudaStreamCaptureStatus st; unsigned long long id; cudaGraph_t g; const cudaGraphNode_t* deps; size_t nd;
    CK(cudaStreamGetCaptureInfo_v2(s, &st, &id, &g, &deps, &nd));
    cudaGraphConditionalHandle hif, hwh;
    CK(cudaGraphConditionalHandleCreate(&hif, g, 0, cudaGraphCondAssignDefault));
    CK(cudaGraphConditionalHandleCreate(&hwh, g, 0, cudaGraphCondAssignDefault));
    set_kernel<<<1, 32, 0, s>>>(d + 8, hif);
    cudaGraph_t body;
    if (add_cond(s, hif, cudaGraphCondTypeIf, &body)) return 1;
    cudaError_t e = cudaStreamBeginCaptureToGraph(b, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) { printf("test2 nested: begin capture failed: %s\n", cudaGetErrorString(e)); }
    else {
      step_kernel<<<1, 32, 0, b>>>(d, 4, hwh, 1);
      cudaGraph_t body2;
      int bad = add_cond(b, hwh, cudaGraphCondTypeWhile, &body2);
      if (!bad) {
        e = cudaStreamBeginCaptureToGraph(b2, body2, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal);
        if (e == cudaSuccess) {
          step_kernel<<<1, 32, 0, b2>>>(d, 4, hwh, 1);
          cudaGraph_t dm; e = cudaStreamEndCapture(b2, &dm);
        }
        if (e != cudaSuccess) printf("test2 inner body: %s\n", cudaGetErrorString(e));
      }
      step_kernel<<<1, 32, 0, b>>>(d + 2, 0, hwh, 0);
      cudaGraph_t dm; e = cudaStreamEndCapture(b, &dm);
      if (e != cudaSuccess) printf("test2 outer body end: %s\n", cudaGetErrorString(e));
    }
    cudaGraph_t graph;
    e = cudaStreamEndCapture(s, &graph);
    if (e != cudaSuccess) printf("test2 end capture: %s\n", cudaGetErrorString(e));
    else {
      cudaGraphExec_t ex;
      e = cudaGraphInstantiate(&ex, graph, 0);
      if (e != cudaSuccess) printf("test2 instantiate: %s\n", cudaGetErrorString(e));
      else {
        CK(cudaGraphLaunch(ex, s)); CK(cudaStreamSynchronize(s));
        int hst[4]; CK(cudaMemcpy(hst, d, 16, cudaMemcpyDeviceToHost));
        printf("test2 nested IF->WHILE: counter %d (want 4) after %d (want 1)\n", hst[0], hst[2]);
        int zero = 0; CK(cudaMemcpy(d + 8, &zero, 4, cudaMemcpyHostToDevice)); CK(cudaMemset(d, 0, 16));
        CK(cudaGraphLaunch(ex, s)); CK(cudaStreamSynchronize(s));
        CK(cudaMemcpy(hst, d, 16, cudaMemcpyDeviceToHost));
        printf("test2 IF false: counter %d (want 0) after %d (want 0)\n", hst[0], hst[2]);
      }
    }
    cudaGetLastError();
  }
  // ---- test 3: cost of empty kernels and of a never-taken WHILE node inside a graph
  {
    const int NK = 2000;
    for (int variant = 0; variant < 3; ++variant) {
      CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
      cudaStreamCaptureStatus st; unsigned long long id; cudaGraph_t g; const cudaGraphNode_t* deps; size_t nd;
      CK(cudaStreamGetCaptureInfo_v2(s, &st, &id, &g, &deps, &nd));
      for (int k = 0; k < NK; ++k) {
        if (variant == 0) empty_kernel<<<1, 32, 0, s>>>(d + 9);
        else if (variant == 1) empty_kernel<<<148, 256, 0, s>>>(d + 9);
        else {
          cudaGraphConditionalHandle h;
          CK(cudaGraphConditionalHandleCreate(&h, g, 0, cudaGraphCondAssignDefault));
          cudaGraph_t body;
          if (add_cond(s, h, cudaGraphCondTypeWhile, &body)) return 1;
          CK(cudaStreamBeginCaptureToGraph(b, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
          step_kernel<<<1, 32, 0, b>>>(d, 0, h, 1);
          cudaGraph_t dm; CK(cudaStreamEndCapture(b, &dm));
        }
      }
      cudaGraph_t graph; CK(cudaStreamEndCapture(s, &graph));
      cudaGraphExec_t ex;
      auto t0 = std::chrono::steady_clock::now();
      CK(cudaGraphInstantiate(&ex, graph, 0));
      auto t1 = std::chrono::steady_clock::now();
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      CK(cudaGraphLaunch(ex, s)); CK(cudaStreamSynchronize(s));
      CK(cudaEventRecord(e0, s)); CK(cudaGraphLaunch(ex, s)); CK(cudaEventRecord(e1, s)); CK(cudaStreamSynchronize(s));
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      printf("test3 variant %d (%s): %d nodes, instantiate %.1f ms, run %.3f ms = %.2f us per node\n", variant,
             variant == 0 ? "empty 1x32" : variant == 1 ? "empty 148x256" : "untaken WHILE", NK,
             std::chrono::duration<double, std::milli>(t1 - t0).count(), ms, 1e3 * ms / NK);
      cudaGraphExecDestroy(ex); cudaGraphDestroy(graph);
    }
    // plain stream launches for comparison
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int k = 0; k < 100; ++k) empty_kernel<<<1, 32, 0, s>>>(d + 9);
    CK(cudaStreamSynchronize(s));
    CK(cudaEventRecord(e0, s));
    for (int k = 0; k < NK; ++k) empty_kernel<<<1, 32, 0, s>>>(d + 9);
    CK(cudaEventRecord(e1, s)); CK(cudaStreamSynchronize(s));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("test3 plain stream: %.2f us per launch\n", 1e3 * ms / NK);
  }
  // ---- test 4: six graphs with WHILE nodes launched on six streams concurrently
  {
    cudaStream_t ss[6]; cudaGraphExec_t ex[6];
    for (int i = 0; i < 6; ++i) {
      CK(cudaStreamCreateWithFlags(&ss[i], cudaStreamNonBlocking));
      CK(cudaStreamBeginCapture(ss[i], cudaStreamCaptureModeThreadLocal));
      cudaStreamCaptureStatus st; unsigned long long id; cudaGraph_t g; const cudaGraphNode_t* deps; size_t nd;
      CK(cudaStreamGetCaptureInfo_v2(ss[i], &st, &id, &g, &deps, &nd));
      for (int k = 0; k < 50; ++k) {
        cudaGraphConditionalHandle h;
        CK(cudaGraphConditionalHandleCreate(&h, g, 0, cudaGraphCondAssignDefault));
        CK(cudaMemsetAsync(d + 16 + i, 0, 4, ss[i]));
        step_kernel<<<1, 32, 0, ss[i]>>>(d + 16 + i, 3, h, 1);
        cudaGraph_t body;
        if (add_cond(ss[i], h, cudaGraphCondTypeWhile, &body)) return 1;
        CK(cudaStreamBeginCaptureToGraph(b, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
        for (int q = 0; q < 10; ++q) work_kernel<<<16, 256, 0, b>>>(x + i * 8192, 4096);
        step_kernel<<<1, 32, 0, b>>>(d + 16 + i, 3, h, 1);
        step_kernel<<<1, 32, 0, b>>>(d + 32 + i, 0, h, 0);
        cudaGraph_t dm; CK(cudaStreamEndCapture(b, &dm));
      }
      cudaGraph_t graph; CK(cudaStreamEndCapture(ss[i], &graph));
      CK(cudaGraphInstantiate(&ex[i], graph, 0));
    }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaMemset(d + 32, 0, 6 * 4));
      CK(cudaDeviceSynchronize());
      auto t0 = std::chrono::steady_clock::now();
      for (int i = 0; i < 6; ++i) CK(cudaGraphLaunch(ex[i], ss[i]));
      CK(cudaDeviceSynchronize());
      auto t1 = std::chrono::steady_clock::now();
      int hst[6]; CK(cudaMemcpy(hst, d + 32, 24, cudaMemcpyDeviceToHost));
      printf("test4 six concurrent graphs: body executions per graph %d %d %d %d %d %d (want 100 each), wall %.3f ms\n", hst[0], hst[1], hst[2], hst[3], hst[4], hst[5],
             std::chrono::duration<double, std::milli>(t1 - t0).count());
    }
    CK(cudaMemset(d + 32, 0, 6 * 4));
    auto t0 = std::chrono::steady_clock::now();
    CK(cudaGraphLaunch(ex[0], ss[0]));
    CK(cudaDeviceSynchronize());
    auto t1 = std::chrono::steady_clock::now();
    printf("test4 one graph alone: wall %.3f ms\n", std::chrono::duration<double, std::milli>(t1 - t0).count());
  }
  printf("done\n");
  return 0;
}
