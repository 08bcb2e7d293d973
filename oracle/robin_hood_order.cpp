// robin_hood_order — TEST INFRASTRUCTURE.  Prints the iteration order of the reference's robin_hood::unordered_map<char, int>
// (include/robin_hood.h, third-party header of the reference, used where it lies) for every subset and insertion order of
// the base letters: the order findconseq (include/VariationUtils.h:403-509) visits the bases of a soft-clip column in.
//   g++ -std=gnu++11 -include limits -w -I/root/reference/include oracle/robin_hood_order.cpp -o oracle/_ref/robin_hood_order
#include <cstdint>
#include <cstddef>
#include <cstring>
#include <limits>
#include <map>
#include <string>
#include <vector>
#include <algorithm>
#include <stdio.h>
using namespace std;
#include "robin_hood.h"
int main(){
  const char keys[5]={'A','C','G','T','N'};
  // every non-empty subset, every insertion order
  for(int mask=1;mask<32;++mask){
    vector<char> sub; for(int i=0;i<5;++i) if(mask&(1<<i)) sub.push_back(keys[i]);
    sort(sub.begin(),sub.end());
    std::string first; bool same=true;
    do{
      robin_hood::unordered_map<char,int> m;
      for(char c:sub) m[c]++;
      robin_hood::unordered_map<char,int> cp = m;
      std::string o; for(auto&e:cp) o.push_back(e.first);
      if(first.empty()) first=o; else if(o!=first) same=false;
    }while(next_permutation(sub.begin(),sub.end()));
    printf("%s -> %s %s\n", std::string(sub.begin(),sub.end()).c_str(), first.c_str(), same?"":"(ORDER DEPENDS ON INSERTION)");
  }
}
