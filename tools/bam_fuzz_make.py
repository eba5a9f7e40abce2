"""Writes /tmp/mut/m0.bam .. m299.bam for tools/bam_fuzz.cpp: a synthetic BAM file whose uncompressed content is damaged and
re-compressed into well-formed BGZF blocks.   python tools/bam_fuzz_make.py [seed]"""
import sys, os, random, zlib, struct
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
import bam_writer as bw
world = bw.synthetic_world(20, config=3, n_samples=1)
os.makedirs('/tmp/n3s', exist_ok=True)
data = open(bw.write_world(world, '/tmp/n3s')[0], 'rb').read()
# inflate all blocks
out=b""; p=0
while p < len(data):
    xlen=struct.unpack_from("<H",data,p+10)[0]
    bsize=None; q=0
    while q+4<=xlen:
        si1,si2,slen=data[p+12+q],data[p+13+q],struct.unpack_from("<H",data,p+14+q)[0]
        if si1==66 and si2==67: bsize=struct.unpack_from("<H",data,p+16+q)[0]+1
        q+=4+slen
    payload=data[p+12+xlen:p+bsize-8]
    out+=zlib.decompress(payload,-15)
    p+=bsize
print("uncompressed", len(out))
rng=random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 5)
os.makedirs('/tmp/mut',exist_ok=True)
def bgzf(raw):
    res=b""
    for k in range(0,len(raw),60000):
        chunk=raw[k:k+60000]
        c=zlib.compressobj(6,zlib.DEFLATED,-15); comp=c.compress(chunk)+c.flush()
        hdr=struct.pack("<BBBBIBBHBBHH",31,139,8,4,0,0,255,6,66,67,2,len(comp)+25)
        res+=hdr+comp+struct.pack("<II",zlib.crc32(chunk)&0xffffffff,len(chunk))
    res+=bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
    return res
for i in range(300):
    b=bytearray(out)
    mode=rng.randrange(3)
    n=1+rng.randrange(3 if mode==0 else 30)
    lo = 0 if mode==2 else 200   # mode 2 may hit the header too
    for _ in range(n):
        pos=rng.randrange(lo,len(b))
        if rng.random()<0.5: b[pos]=rng.randrange(256)
        else: struct.pack_into("<I", b, min(pos,len(b)-4), rng.choice([0,1,0xffffffff,0x7fffffff,0x80000000,rng.randrange(1<<32)]))
    open('/tmp/mut/m%d.bam'%i,'wb').write(bgzf(bytes(b)))
print("written")
