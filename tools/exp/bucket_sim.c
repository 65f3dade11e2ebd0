/* bucket_sim.c -- CPU simulation of the encoder's bucket dictionary (128 buckets of 32 slots,
 * bucket = (q(prefix) & 127) ^ hash7(byte), a full bucket overflows into the next one) on config-3
 * strips: bucket loads per input byte for a scramble q(code) = code * M + D mod 4096 and a byte
 * hash ((byte * BM) >> BS) & 127.  Used to choose M = 0xC55, D = -255 * M, BM = 0x83F, BS = 2
 * (encode_kernels.cu): 1.0745 loads per byte against 1.0734 for round 1's 0x9E5 / 0x6A7.
 * Build: gcc -O2 -o bucket_sim bucket_sim.c ../wlgen/wlgen.c -pthread
 * Run:   ./bucket_sim <strips> <M> <BM> <BS> [D]        e.g. ./bucket_sim 1024 0xC55 0x83F 2 0x755 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
uint64_t wl_tiff_strip_len(uint64_t, uint64_t, uint64_t, uint64_t);
void wl_tiff_strips_fill(uint64_t, uint64_t, uint64_t, uint64_t, uint64_t, const uint64_t*, uint8_t*);
static uint32_t M, BM, BS;
static uint32_t D; static uint32_t q(uint32_t c){ return (c*M + D) & 0xFFF; }
static uint32_t h7(uint32_t k){ return ((k*BM) >> BS) & 0x7F; }
static uint32_t tbl[128][32]; static int cnt[128];
int main(int argc,char**argv){
  uint64_t n = strtoull(argv[1],0,0); M = strtoul(argv[2],0,0); BM = strtoul(argv[3],0,0); BS = strtoul(argv[4],0,0); D = argc > 5 ? strtoul(argv[5],0,0) : 0xFFF;
  uint64_t seed = 0x5A172E60ull + 3;
  uint64_t* off = malloc((n+1)*8); off[0]=0;
  for (uint64_t i=0;i<n;i++) off[i+1]=off[i]+wl_tiff_strip_len(seed,i,8192,57345);
  uint8_t* buf = malloc(off[n]); wl_tiff_strips_fill(seed,0,n,8192,57345,off,buf);
  uint64_t loads[4]={0}, bytes[4]={0}, maxchain=0;
  for (uint64_t s=0;s<n;s++){
    const uint8_t* p = buf+off[s]; uint64_t len = off[s+1]-off[s];
    memset(cnt,0,sizeof cnt); uint32_t next=258; uint32_t prefix=p[0];
    for (uint64_t i=1;i<len;i++){
      uint32_t k=p[i]; uint32_t key=(q(prefix)<<8)|k; uint32_t b=(q(prefix)&0x7F)^h7(k); uint64_t ch=0;
      bytes[s&3]++;
      for(;;){
        loads[s&3]++; ch++;
        int found=-1; for(int j=0;j<cnt[b];j++) if((tbl[b][j]>>12)==key){found=j;break;}
        if(found>=0){prefix=tbl[b][found]&0xFFF; break;}
        if(cnt[b]<32){ tbl[b][cnt[b]++]=(key<<12)|next; next++; prefix=k;
          if(next==4094+1){ /* table reset (approx: clear at 4094 entries) */ memset(cnt,0,sizeof cnt); next=258; }
          break; }
        b=(b+1)&0x7F;
      }
      if(ch>maxchain)maxchain=ch;
    }
  }
  for(int c=0;c<4;c++) printf("class %d: %.4f loads/byte\n",c,(double)loads[c]/bytes[c]);
  printf("all: %.4f  maxchain %llu\n",(double)(loads[0]+loads[1]+loads[2]+loads[3])/(bytes[0]+bytes[1]+bytes[2]+bytes[3]),(unsigned long long)maxchain);
}
