/* Test infrastructure: cpic_b200/csrc/host/glibc_rand.h (the recurrence the device initialiser draws
 * from, and its jump-ahead) against the C library's own srand()/rand(). Built and run by tests/test_host.py. */
#include "glibc_rand.h"
#include <stdio.h>
#include <stdlib.h>
int main(){
	static GlibcRandJump J;
	int bad=0;
	unsigned seeds[]={0,1,138,139,4242424242u,2147483647u};
	for(unsigned s: seeds){
		uint32_t st[31]; glibc_rand_seed(s, st);
		srand(s);
		for(int k=0;k<5000;k++){ uint32_t a=glibc_rand_next(st); int b=rand(); if((int)a!=b){ if(bad<5) printf("seed %u k %d: %u vs %d\n", s,k,a,b); bad++; } }
		// jump test
		uint32_t s2[31]; glibc_rand_seed(s, s2); J.jump(s2, 123457);
		srand(s); for(int k=0;k<123457;k++) rand();
		for(int k=0;k<100;k++){ uint32_t a=glibc_rand_next(s2); int b=rand(); if((int)a!=b) bad++; }
		uint32_t s3[31]; glibc_rand_seed(s, s3); uint32_t m[31][31]; J.power(1000,m);
		for(int t=0;t<7;t++) GlibcRandJump::apply(m,s3);
		srand(s); for(int k=0;k<7000;k++) rand();
		for(int k=0;k<100;k++){ uint32_t a=glibc_rand_next(s3); int b=rand(); if((int)a!=b) bad++; }
	}
	printf("bad=%d\n",bad); return bad!=0;
}
