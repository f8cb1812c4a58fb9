#include "fastq_frame.hpp"
#include "gzip_inflate.hpp"
#include "genotype_host.hpp"
#include <fcntl.h>
#include <unistd.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
using namespace drprg;
int main(int argc,char**argv){
  for (int a=1;a<argc;++a){
    std::string path=argv[a];
    int fd=open(path.c_str(),O_RDONLY); size_t size=lseek(fd,0,SEEK_END);
    if (path.size()>3 && path.substr(path.size()-3)==".gz"){
      std::vector<uint8_t> z(size+64,0); pread(fd,z.data(),size,0);
      char* text=nullptr; size_t tn=0;
      bool ok=parallel_gunzip(z.data(),size,8,&text,&tn);
      printf("%s gunzip ok=%d n=%zu\n",path.c_str(),ok,tn);
      if(ok){ // frame from memory with an exact-size buffer (ASAN catches over-reads)
        char* exact=(char*)malloc(tn); memcpy(exact,text,tn); free(text);
        TextSource src; src.mem=exact; src.size=tn; std::vector<char> buf(tn/2+64); std::vector<FramedSlice> sl;
        bool f=fastq_frame_text(src,8,buf.data(),sl); size_t n=0; for(auto&z2:sl)n+=z2.st.n_reads;
        printf("  framed ok=%d reads=%zu\n",f,n); free(exact);
      }
    } else {
      TextSource src; src.fd=fd; src.size=size; std::vector<char> buf(size/2+64); std::vector<FramedSlice> sl;
      bool f=fastq_frame_text(src,8,buf.data(),sl); size_t n=0; for(auto&z2:sl)n+=z2.st.n_reads;
      printf("%s framed ok=%d reads=%zu slices=%zu\n",path.c_str(),f,n,sl.size());
      // and from memory, exact-size allocation
      char* exact=(char*)malloc(size); pread(fd,exact,size,0);
      TextSource m; m.mem=exact; m.size=size; std::vector<FramedSlice> sl2; std::vector<char> buf2(size/2+64);
      bool f2=fastq_frame_text(m,8,buf2.data(),sl2); size_t n2=0; for(auto&z2:sl2)n2+=z2.st.n_reads;
      printf("  from memory ok=%d reads=%zu\n",f2,n2); free(exact);
    }
    close(fd);
  }
}
