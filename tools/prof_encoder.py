import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from daclip_b200 import synthetic
from daclip_b200.daclip import DaCLIP
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
clip = DaCLIP().load_reference_state_dict(synthetic.daclip_visual_state_dict(10)).to(dev).eval()
img = torch.randn(B, 3, 224, 224, device=dev)
for _ in range(3):
    clip.encode_image(img, control=True)
torch.cuda.synchronize()
