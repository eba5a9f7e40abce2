// Host check of csrc/host/poa.cpp's edit distances (Myers bit-vector recurrence on 64-bit words, and its cut-off version that
// evaluates only the blocks within k diagonals) against the plain cell recurrence on seeded pairs: random strings, noisy
// copies, copies with a block removed; cut-offs 0, 1, random, the distance itself and the distance minus one.
// Built and run by tests/test_assembly.py::test_host_edit_distances_fuzz.
#include <cstdio>
#include <random>
#include <string>
#include <vector>
#include <algorithm>
#include "poa.h"
static int plain(const std::string& a, const std::string& b){ int n=a.size(),m=b.size(); std::vector<int> p(m+1),c(m+1); for(int j=0;j<=m;++j)p[j]=j; for(int i=1;i<=n;++i){c[0]=i; for(int j=1;j<=m;++j) c[j]=std::min(std::min(p[j]+1,c[j-1]+1),p[j-1]+(a[i-1]!=b[j-1])); std::swap(p,c);} return p[m]; }
int main(){ std::mt19937 g(5); long bad=0, n_exact=0, total=0;
 for(int it=0; it<12000; ++it){ int n=g()%400; if(it%50==0) n=g()%1500; std::string a; for(int i=0;i<n;++i)a.push_back("ACGT"[g()%4]); std::string b;
   double er = (g()%5)*0.03; if (it%7==0) { int m=g()%400; for(int i=0;i<m;++i)b.push_back("ACGT"[g()%4]); } else { for(char ch: a){ double u=(g()%10000)/10000.0; if(u<er/3) continue; if(u<2*er/3){b.push_back("ACGT"[g()%4]); continue;} if(u<er){b.push_back(ch); b.push_back("ACGT"[g()%4]); continue;} b.push_back(ch);} if(it%11==0 && b.size()>40){ int p=g()%(b.size()-30); b.erase(p, g()%30);} }
   int ed=plain(a,b); if (ltr::edit_distance(a,b)!=ed && !(a.empty()||b.empty())) {++bad; if(bad<5) printf("full mismatch %d\n",it);}
   int ks[6]={0,1,(int)(g()%70),(int)(g()%300),ed,std::max(0,ed-1)};
   for(int k: ks){ int r=ltr::bounded_edit_distance(a,b,k); int want= ed<=k?ed:k+1; ++total; if(r!=want){++bad; if(bad<10) printf("it %d n %zu m %zu k %d got %d want %d (ed %d)\n",it,a.size(),b.size(),k,r,want,ed);} else if(ed<=k) ++n_exact; }
 }
 printf("checks %ld exact-range %ld bad %ld\n", total, n_exact, bad); return bad!=0; }
