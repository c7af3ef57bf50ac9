mkdir -p gpurun_out
timeout 300 python -c "
import ctypes as C, sys
sys.path.insert(0,'.')
from generative_recommenders_b200 import _lib
buf=C.create_string_buffer(1<<16)
r=_lib.lib().hstu_umma_selftest(buf,len(buf))
open('gpurun_out/selftest.txt','w').write(buf.value.decode()+'\nrc=%d\n'%r)
print(buf.value.decode()[-2600:], r)
"
