#!/bin/bash
# registers / spills of every kernel in kernels.cu (nvcc -Xptxas -v), demangled
/usr/local/cuda/bin/nvcc -ccbin /usr/bin/g++ -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC,-ffp-contract=off $EXTRA -Iinclude -Iraytracer-2025_b200/csrc -Xptxas -v -c raytracer-2025_b200/csrc/kernels.cu -o /tmp/k.o 2>&1 | python3 -c "
import sys,re,subprocess
name=None; spill=''
for line in sys.stdin:
    m=re.search(r\"Compiling entry function '(\S+)'\", line)
    if m: name=m.group(1); continue
    if 'spill' in line: spill=line.strip().replace('ptxas info    : ','')
    m=re.search(r'Used (\d+) registers', line)
    if m and name:
        dm=subprocess.run(['c++filt',name],capture_output=True,text=True).stdout.strip().split('(')[0].replace('void rt::','')
        print(f'{dm:50s} regs {m.group(1):>3s}  {spill}')
        name=None
" | sort
