# scratch (not committed): first timing run with an oracle-built template until GPU map-gen lands
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, 'oracle'); sys.path.insert(0,'tests')
import bench, oracle as O
def mk(vr, cfg, width, device):
    ot = O.build_template(cfg, width)
    for d in ot.inputs: d['vignette'] = None
    return vr.MapperTemplate.from_arrays(ot.out_size, ot.inputs, ot.seam_masks)
bench.make_template = mk
bench.main()
