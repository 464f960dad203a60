"""Opcode histogram of a kernel in an object file (cuobjdump -sass), restricted to the region around its tensor
instructions (HMMA): python tools/sass_histogram.py build/attn_f16.o "Li16ELi13ELi72E"

A static count (unrolled code), used to compare kernel variants before spending GPU time."""
import sys,re,subprocess,collections
obj,pat=sys.argv[1],sys.argv[2]
out=subprocess.run(['cuobjdump','-sass',obj],capture_output=True,text=True).stdout
fn=None;ins=[]
for line in out.splitlines():
    m=re.search(r'Function : (\S+)',line)
    if m: fn=m.group(1);continue
    m=re.match(r'\s+/\*[0-9a-f]{4,6}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)(\S*)',line)
    if m and fn and re.search(pat,fn): ins.append(m.group(2)+m.group(3).split(' ')[0])
first=next(i for i,x in enumerate(ins) if x.startswith('HMMA'))
last=max(i for i,x in enumerate(ins) if x.startswith('HMMA'))
body=ins[first-60:last+80]
c=collections.Counter(re.split(r'[.;]',x)[0] for x in body)
print('body',len(body),'pre',first)
print(' '.join(f'{k}={v}' for k,v in c.most_common(40)))
c2=collections.Counter(x.rstrip(';') for x in body if x.startswith(('ISETP','IMAD','FMUL','MOV','LOP3')))
print(' '.join(f'{k}={v}' for k,v in c2.most_common(24)))
