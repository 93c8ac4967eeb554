#include <nvrtc.h>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>
static bool compile(const std::string& src, bool minimal, std::vector<char>* cubin) {
  std::vector<const char*> o = {"--gpu-architecture=sm_100a", "--fmad=false", "--prec-div=true", "--prec-sqrt=true", "--ftz=false", "-lineinfo", "--std=c++17", "-default-device"};
  if (minimal) o.push_back("--minimal");
  nvrtcProgram p; nvrtcCreateProgram(&p, src.c_str(), "vkjit_trace.cu", 0, nullptr, nullptr);
  nvrtcResult r = nvrtcCompileProgram(p, (int)o.size(), o.data());
  if (r != NVRTC_SUCCESS) { size_t ls=0; nvrtcGetProgramLogSize(p,&ls); std::string log(ls,0); nvrtcGetProgramLog(p,&log[0]); fprintf(stderr,"%s\n",log.c_str()); return false; }
  size_t cs=0; nvrtcGetCUBINSize(p,&cs); cubin->resize(cs); nvrtcGetCUBIN(p,cubin->data()); nvrtcDestroyProgram(&p); return true;
}
int main(int argc, char** argv) {
  int bad = 0;
  for (int i = 1; i < argc; ++i) {
    std::ifstream f(argv[i]); std::stringstream ss; ss << f.rdbuf(); std::string src = ss.str();
    std::vector<char> a, b;
    bool ok = compile(src, false, &a) && compile(src, true, &b);
    bool same = ok && a.size()==b.size() && memcmp(a.data(), b.data(), a.size())==0;
    printf("%s: %s (%zu bytes)\n", argv[i], !ok ? "COMPILE ERROR" : same ? "identical cubin" : "DIFFERENT", a.size());
    if (!same) ++bad;
  }
  return bad;
}
