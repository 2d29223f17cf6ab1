/* hash_quality.c -- statistical comparison of the pair-hash revisions v2 and v3.  Development tool.
 *
 * For N objects and M = 1024 nodes with hashed seeds it counts, per node, how often the node has the largest u
 * (winner), the second largest (runner-up: where a leaving node's objects go), and for the first 8 nodes the
 * runner-up conditioned on that node winning.  Equal weights, so the rendezvous rule is "largest u".  Prints the
 * three chi-squares against the uniform expectation (df and its standard deviation beside them) and max/min load.
 *   gcc -O2 -pthread -o tools/hash_quality tools/hash_quality.c -lm && tools/hash_quality 100000000 8
 * Result recorded in DESIGN.md 3.3: v2 and v3 ("Z", two multiply-adds) are indistinguishable at 1e8 objects.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <pthread.h>
static inline uint64_t mix64(uint64_t x){x^=x>>30;x*=0xBF58476D1CE4E5B9ull;x^=x>>27;x*=0x94D049BB133111EBull;x^=x>>31;return x;}
#define M 1024
static uint32_t s0[M],s1[M],s2[M];
typedef struct { int variant; long lo,hi; long *cnt; long *cnt2; long *cond; } job;
static void* work(void*p){ job*j=(job*)p;
 for(long i=j->lo;i<j->hi;i++){ uint64_t key=mix64(0x9E3779B97F4A7C15ull*(i+1)^1); uint64_t h=mix64(key^0xD6E8FEB86659FD93ull); uint32_t a=(uint32_t)h,b=(uint32_t)(h>>32)|1u,ab=a*b;
  uint32_t best=0,sec=0; int bj=-1,sj=-1;
  for(int k=0;k<M;k++){ uint32_t pp=s0[k]*b+ab; uint32_t u;
    if(j->variant==0){ uint32_t q=pp^(pp>>15)^s1[k]; u=q*0x9E3779B1u+s2[k]; }
    else { u=pp*(s1[k]|1u)+s2[k]; }
    if(bj<0||u>best){sec=best;sj=bj;best=u;bj=k;} else if(sj<0||u>sec){sec=u;sj=k;} }
  j->cnt[bj]++; j->cnt2[sj]++; if(bj<8) j->cond[bj*M+sj]++; }
 return 0;}
int main(int argc,char**argv){ long N=atol(argv[1]); int T=atoi(argv[2]);
 for(int k=0;k<M;k++){ uint64_t s=mix64((uint64_t)k*0x100000001b3ull+12345); s0[k]=(uint32_t)s; s1[k]=(uint32_t)(s>>32); s2[k]=(uint32_t)mix64(s^0xA0761D6478BD642Full);} 
 for(int v=0;v<2;v++){ pthread_t th[64]; job jobs[64]; long *cnt=calloc(M*T,sizeof(long)),*cnt2=calloc(M*T,sizeof(long)),*cond=calloc(8L*M*T,sizeof(long));
  for(int t=0;t<T;t++){ jobs[t].variant=v; jobs[t].lo=N*t/T; jobs[t].hi=N*(t+1)/T; jobs[t].cnt=cnt+M*t; jobs[t].cnt2=cnt2+M*t; jobs[t].cond=cond+8L*M*t; pthread_create(&th[t],0,work,&jobs[t]); }
  for(int t=0;t<T;t++) pthread_join(th[t],0);
  double e=(double)N/M,chi=0,chi2=0; long mx=0,mn=1L<<60;
  for(int k=0;k<M;k++){ long c=0,c2=0; for(int t=0;t<T;t++){c+=cnt[M*t+k];c2+=cnt2[M*t+k];} chi+=(c-e)*(c-e)/e; chi2+=(c2-e)*(c2-e)/e; if(c>mx)mx=c; if(c<mn)mn=c; }
  double cc=0; long df=0; for(int w=0;w<8;w++){ long n=0; for(int k=0;k<M;k++){ long c=0; for(int t=0;t<T;t++) c+=cond[8L*M*t+w*M+k]; n+=c; }
     double ee=(double)n/(M-1); for(int k=0;k<M;k++){ if(k==w) continue; long c=0; for(int t=0;t<T;t++) c+=cond[8L*M*t+w*M+k]; cc+=(c-ee)*(c-ee)/ee; } df+=M-2; }
  printf("variant %s N=%ld: chi_win=%.0f chi_sec=%.0f (df %d sd %.0f) cond=%.0f (df %ld sd %.0f) max/min load %.4f %.4f\n", v?"Z (2 IMAD)":"v2", N, chi, chi2, M-1, sqrt(2.0*(M-1)), cc, df, sqrt(2.0*df), mx/e, mn/e);
  free(cnt);free(cnt2);free(cond);} }
