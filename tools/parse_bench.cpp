// Reader-only benchmark: ParallelReader over a file, no GPU. usage: parse_bench FILE THREADS RANGE_MB [q]
// build: g++ -O3 -std=c++17 -pthread -Ibgreat_b200/csrc -o /tmp/parse_bench tools/parse_bench.cpp bgreat_b200/csrc/bgr_host.cpp
#include "bgr_host.h"
#include <chrono>
#include <stdio.h>
using namespace bgr;
int main(int argc,char**argv){
  MappedFile mf; if(!mf.open(argv[1])) return 1;
  unsigned T=atoi(argv[2]); size_t range=(size_t)atoll(argv[3])<<20; bool fq=argc>4;
  for(int rep=0;rep<3;++rep){
  auto t0=std::chrono::steady_clock::now();
  ParallelReader pr(mf.data,mf.size,31,fq,T,range);
  ParsedBatch pb; size_t n=0,c=0;
  do{ pb.clear(); pr.next(pb); n+=pb.size(); c+=pb.seqs.n;}while(!pr.done());
  double dt=std::chrono::duration<double>(std::chrono::steady_clock::now()-t0).count();
  printf("T=%u reads=%zu chars=%zu %.3f s %.2f GB/s %.1f Mreads/s\n",T,n,c,dt,mf.size/dt/1e9,n/dt/1e6);
  }
}
