import ctypes as C, sys, time
sys.path.insert(0,'/root/repo')
import torch
from handwritten_math_ocr_api_b200 import _lib, FormulaRecognitionModel
lib=_lib.load(); n=C.c_int(); torch.cuda.init(); torch.zeros(1).cuda()
print('rc', lib.hmocr_decode_max_clusters(C.byref(n)), 'max active clusters', n.value, lib.hmocr_last_error())
from handwritten_math_ocr_api_b200.layout import ModelConfig
from handwritten_math_ocr_api_b200.synthetic import synth_state_dict, synth_images
cfg=ModelConfig(); m=FormulaRecognitionModel(cfg.vocab_size); m.load_state_dict(synth_state_dict(cfg, eos_bias_sigma=0.0))
imgs=synth_images(8).cuda()
for B in [1,16,32,64,128,192,208,224,240,256,512]:
    x=imgs.repeat((B+7)//8,1,1,1)[:B].contiguous()
    enc=m.encoder(x)
    for _ in range(2):
        m.generate(encoder_out=enc, max_len=150)
    torch.cuda.synchronize(); t=time.perf_counter(); m.generate(encoder_out=enc, max_len=150); torch.cuda.synchronize(); dt=time.perf_counter()-t
    print(B, 'decode ms', round(dt*1e3,2), 'us/step', round(dt*1e6/150,1))
