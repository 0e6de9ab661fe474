/* noop_runs.c -- TEST / ANALYSIS INFRASTRUCTURE (not part of the product): the oracle's plane grower
 * (oracle/bseg_oracle.c orc_grow, i.e. my_function.cpp:180-258) instrumented to record, for every Broad() call of the
 * transactions that pass depth 0, how many points it accepted and whether any neighbour was still free.  DESIGN.md 6
 * quotes its statistics (three calls out of four accept nothing, 88 % of those in runs of 32 and more).
 * Build + run: python oracle/tools/noop_runs.py
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "../../buildingsegment_b200/csrc/bseg_arith.h"
typedef struct { int64_t cur, end; } fr;
// records accept counts per call for transactions that pass depth 0; out: cnts (uint8), tx_off
int64_t sim_grow(const int32_t* xyz, int64_t n, const double* normal, const int32_t* neigh, int K,
                 int th_thickness, int th_count, double th_dot, uint8_t* cnts, int64_t cap, int64_t* tx_off, int64_t cap_tx, int64_t* ntx_out)
{
  int32_t* plane_idx = malloc(n*4); for (int64_t i=0;i<n;++i) plane_idx[i]=-1;
  int64_t cap_list=1<<20, cap_stack=1<<20; int32_t* list=malloc(cap_list*4); fr* stack=malloc(cap_stack*sizeof(fr));
  int64_t nc=0, ntx=0; int32_t cur_id=1; tx_off[0]=0;
  for (int64_t i=0;i<n;++i){
    if (plane_idx[i]!=-1) continue;
    int64_t len=0, sp=0; list[len++]=(int32_t)i;
    double mn0=normal[3*i],mn1=normal[3*i+1],mn2=normal[3*i+2];
    int32_t mc0=xyz[3*i],mc1=xyz[3*i+1],mc2=xyz[3*i+2];
    double sn0=0.0+mn0,sn1=0.0+mn1,sn2=0.0+mn2; uint32_t sc0=mc0,sc1=mc1,sc2=mc2;
    int64_t node=i; int depth0=1, ok=1; int64_t c0=nc;
    for(;;){
      int64_t s0=len; int ncand=0;
      for(int j=1;j<K;++j){ int32_t id=neigh[node*K+j]; if(id<0)continue;
        if(plane_idx[id]<=0){ ++ncand;
          int32_t p0=(int32_t)((uint32_t)xyz[3*id]-(uint32_t)mc0),p1=(int32_t)((uint32_t)xyz[3*id+1]-(uint32_t)mc1),p2=(int32_t)((uint32_t)xyz[3*id+2]-(uint32_t)mc2);
          double dist=fabs(p0*mn0+p1*mn1+p2*mn2);
          if(dist<=(double)th_thickness && mn0*normal[3*id]+mn1*normal[3*id+1]+mn2*normal[3*id+2]>=th_dot){
            if(len==cap_list){cap_list*=2;list=realloc(list,cap_list*4);}
            list[len++]=id; plane_idx[id]=cur_id;
            sn0+=normal[3*id];sn1+=normal[3*id+1];sn2+=normal[3*id+2];
            sc0+=(uint32_t)xyz[3*id];sc1+=(uint32_t)xyz[3*id+1];sc2+=(uint32_t)xyz[3*id+2];
          }}}
      if(depth0 && (len-s0)<K-1){ok=0;break;}
      depth0=0;
      if(nc<cap) cnts[nc]=(uint8_t)((len-s0) | ((ncand>0)?0x80:0) | ((ncand>3)?0x40:0)); ++nc;
      { double nn=bseg_sqrt((sn0*sn0)+(sn1*sn1)+(sn2*sn2)); mn0=sn0/nn;mn1=sn1/nn;mn2=sn2/nn; uint64_t dv=len;
        mc0=(int32_t)(uint32_t)(((uint64_t)(int64_t)(int32_t)sc0)/dv);mc1=(int32_t)(uint32_t)(((uint64_t)(int64_t)(int32_t)sc1)/dv);mc2=(int32_t)(uint32_t)(((uint64_t)(int64_t)(int32_t)sc2)/dv);}
      if(len>s0){ if(sp==cap_stack){cap_stack*=2;stack=realloc(stack,cap_stack*sizeof(fr));} stack[sp].cur=s0;stack[sp].end=len;++sp;}
      while(sp>0&&stack[sp-1].cur==stack[sp-1].end)--sp;
      if(sp==0)break;
      node=list[stack[sp-1].cur++];
    }
    if(!ok){ continue; }
    if(ntx<cap_tx){ tx_off[ntx+1]=nc; } ++ntx;
    if(len>th_count){ ++cur_id; } else { for(int64_t t=0;t<len;++t)plane_idx[list[t]]=-1; }
  }
  *ntx_out=ntx; return nc;
}
