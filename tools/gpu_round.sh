python - <<'PY'
import sys, struct, subprocess, os, time
sys.path.insert(0,'tests')
import numpy as np, helpers as H
l,r,pk=H.synth(1,44100*60,16,want_packed=True)
n=pk.size
hdr=b"RIFF"+struct.pack("<I",36+n)+b"WAVEfmt "+struct.pack("<IHHIIHH",16,1,2,44100,44100*4,4,16)
for i in range(3): open(f'/dev/shm/c{i}.wav','wb').write(hdr+b"data"+struct.pack("<I",n)+pk.tobytes())
open('/dev/shm/l.txt','w').write("".join(f"encode /dev/shm/c{i}.wav /dev/shm/c{i}.lac\ndecode /dev/shm/c{i}.lac /dev/shm/c{i}b.wav\n" for i in range(3)))
cli=str(H.PKG_DIR/'host'/'lac_cli')
env=dict(os.environ, LAC_TIMING='1', LACB_TRACE='1')
r=subprocess.run([cli,'batch','/dev/shm/l.txt'],capture_output=True,text=True,env=env); print(r.stderr)
PY
