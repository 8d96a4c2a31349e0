// Measurement aid (not part of the product): how fast a tmpfs output file can be written -- T threads pwrite()-ing
// one file (mode 0), T threads copying into a shared mapping of one file (mode 1), one file per writer thread (mode 2).
// What it shows on the 8-core build box: one file takes 2.4 GB/s by write() however many threads (inode lock) and
// 3.8 GB/s through a mapping with 8 threads (page allocation); only separate files scale. The drop-in command line
// has to produce the reference's two or three files, so its writer is bound there.
//   gcc -O2 -pthread -o out_bench tools/out_bench.c && ./out_bench MB THREADS MODE
#define _GNU_SOURCE
#include <fcntl.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <time.h>
#include <unistd.h>
static double now(){struct timespec t;clock_gettime(CLOCK_MONOTONIC,&t);return t.tv_sec+1e-9*t.tv_nsec;}
static char* src; static size_t CH=16u<<20;
struct A{int fd; size_t off,len; char* map;};
static void* wr(void* p){struct A* a=p; for(size_t o=0;o<a->len;o+=CH){size_t n=a->len-o<CH?a->len-o:CH; if(pwrite(a->fd,src,n,a->off+o)!=(ssize_t)n)perror("pwrite");} return 0;}
static void* mm(void* p){struct A* a=p; for(size_t o=0;o<a->len;o+=CH){size_t n=a->len-o<CH?a->len-o:CH; memcpy(a->map+a->off+o,src,n);} return 0;}
int main(int argc,char**argv){
  size_t total=(size_t)atol(argv[1])<<20; int T=atoi(argv[2]); int mode=atoi(argv[3]);
  src=malloc(CH); memset(src,'A',CH);
  int nf = mode==2?T:1;   // mode 0: pwrite T threads one file; 1: mmap T threads one file; 2: write, one file per thread
  int fds[64]; char name[64];
  for(int i=0;i<nf;i++){sprintf(name,"/dev/shm/wtest_%d",i); fds[i]=open(name,O_RDWR|O_CREAT|O_TRUNC,0666);}
  double t0=now();
  pthread_t th[64]; struct A a[64]; char* map=0;
  if(mode==1){ftruncate(fds[0],total); map=mmap(0,total,PROT_READ|PROT_WRITE,MAP_SHARED,fds[0],0);}
  for(int t=0;t<T;t++){a[t].fd=fds[mode==2?t:0]; a[t].off=mode==2?0:total/T*t; a[t].len=total/T; a[t].map=map; pthread_create(&th[t],0,mode==1?mm:wr,&a[t]);}
  for(int t=0;t<T;t++)pthread_join(th[t],0);
  if(map)munmap(map,total);
  double dt=now()-t0; printf("mode %d T %d: %.2f GB/s\n",mode,T,total/dt/1e9);
  for(int i=0;i<nf;i++){close(fds[i]); sprintf(name,"/dev/shm/wtest_%d",i); unlink(name);}
}
