"""TEST INFRASTRUCTURE (build container only — needs /root/reference): run the reference's UNMODIFIED training script
`train64.py` / `train128.py` for a few iterations over a chosen implementation of the `models` package.

    python tests/ref_harness.py --impl b200-emul|reference --size 64|128 --out DIR [--niter 2] [--batch 3]

--impl reference : `models` resolves to /root/reference/models (the reference's own PyTorch modules, CPU)
--impl b200-emul : `models` resolves to this repository's module surface; with no GPU in the build container the kernel
                   namespace is the CPU emulation of the C ABI (tests/abi_emul.py), exactly as in tests/test_wiring_cpu.py

The script itself is imported from /root/reference and its `main(config)` is called.  What has to be shimmed, and why
(SURVEY.md F6 — none of it touches the path under test):
  * `tensorboardX`, `h5py`, `imageio` are not installed -> empty stub modules (tensorboard logging is switched off);
  * `data.vg_custom_mask.get_dataloader` needs the Visual Genome h5 files and is called with an `image_size=` keyword it
    does not accept (train64.py:91 vs vg_custom_mask.py:224) -> a synthetic loader with the same return contract
    (batch tuple of vg_collate_fn, `.dataset.num_objects`), every object annotated (train64.py:163 breaks on torch >= 1.2
    for un-annotated objects);
  * the hard-coded devices `torch.device('cuda:1')` (train64.py:85) and `map_location="cuda:0"`
    (utils/model_saver_iter.py:40) -> the name `torch` inside those two modules is a proxy that maps both to the CPU;
  * `matrix_obj_vs_att.pt` is loaded from the working directory -> a synthetic co-occurrence matrix is written there.
Initial weights come through the reference's own `load_model` from `iter-0_*.pkl` checkpoints written here from the
oracle's seeded states, so both implementations start identically; the final `iter-N_*.pkl` files the script saves and
the losses it prints are what tests/test_reference_scripts.py compares."""
import argparse
import io
import os
import random
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "attribute-guided-image-generation-from-layout_b200")
REF = "/root/reference"


class TorchProxy(types.ModuleType):
    """`torch` as seen by train64.py / model_saver_iter.py: everything is the real module except the hard-coded devices"""

    def __init__(self, real, device):
        super().__init__("torch")
        self.__dict__["_real"], self.__dict__["_dev"] = real, device

    def __getattr__(self, name):
        return getattr(self.__dict__["_real"], name)

    def device(self, *a, **kw):
        return self.__dict__["_real"].device(self.__dict__["_dev"])

    def load(self, f, map_location=None, **kw):
        return self.__dict__["_real"].load(f, map_location=self.__dict__["_dev"], **kw)


def synthetic_loader_module(image_size, n_batches, seed0):
    import torch
    sys.path.insert(0, ROOT)
    from oracle import gan_oracle as O           # test infrastructure: the synthetic VG-shaped batch generator

    class _Dataset:
        num_objects = O.NUM_OBJECTS

    class _Loader:
        dataset = _Dataset()

        def __init__(self, batch_size):
            self.batch_size = batch_size

        def __iter__(self):
            for i in range(n_batches):
                b = O.synth_batch(self.batch_size, image_size, None, seed0 + i)
                yield (b["imgs"], b["objs"], b["boxes"], b["masks"], b["obj_to_img"], b["attribute"], b["masks_shift"],
                       b["boxes_shift"])

    mod = types.ModuleType("data.vg_custom_mask")

    def get_dataloader(batch_size=10, VG_DIR=None, VG_IMG_DIR=None, attribute_embedding=128, image_size=None):
        return _Loader(batch_size), _Loader(batch_size)

    mod.get_dataloader = get_dataloader
    return mod


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", required=True, choices=["b200-emul", "reference"])
    ap.add_argument("--size", type=int, default=64)
    ap.add_argument("--out", required=True)
    ap.add_argument("--niter", type=int, default=2)
    ap.add_argument("--batch", type=int, default=3)
    ap.add_argument("--script", default="train", choices=["train", "test"],
                    help="train: train64/128.py main(config); test: test64.py (runs its inference / attribute-edit loop on import)")
    args = ap.parse_args()
    import torch
    torch.set_num_threads(max(1, os.cpu_count() or 1))

    written = {}
    for name in ("tensorboardX", "h5py", "imageio"):
        m = types.ModuleType(name)
        if name == "tensorboardX":
            m.SummaryWriter = lambda *a, **k: None
        if name == "imageio":                      # test64.py:13 — images the script "writes" are recorded instead
            m.imwrite = lambda path, arr: written.__setitem__(os.path.basename(path), torch.as_tensor(arr).clone())
        sys.modules[name] = m
    # `models` first from the implementation under test, everything else (utils/, data/, attribute_*.py) from the reference
    if args.impl == "b200-emul":
        sys.path[:0] = [PKG, os.path.join(ROOT, "tests"), ROOT]
        from b200gan import _lib
        if not torch.cuda.is_available():
            from abi_emul import EmulKernels
            _lib.K = EmulKernels()
    sys.path.append(REF)
    sys.path.insert(0, ROOT)
    from oracle import gan_oracle as O
    sys.modules["data.vg_custom_mask"] = synthetic_loader_module(args.size, args.niter, 500)

    os.makedirs(args.out, exist_ok=True)
    os.chdir(args.out)
    g = torch.Generator().manual_seed(0)
    matrix = torch.randint(0, 5000, (O.NUM_OBJECTS, O.NUM_ATTRIBUTES), generator=g).float()
    matrix[0] = 0
    torch.save(matrix, "matrix_obj_vs_att.pt")

    import utils.model_saver_iter as saver
    dev = "cuda:0" if (torch.cuda.is_available() and args.impl == "b200-emul") else "cpu"
    proxy = TorchProxy(torch, dev)
    saver.torch = proxy
    states = O.make_states(args.size, 0)

    def check_models():
        import models
        expect = PKG if args.impl == "b200-emul" else REF
        assert os.path.abspath(models.__file__).startswith(expect), (models.__file__, expect)

    if args.script == "test":
        return run_test_script(args, torch, proxy, states, written, check_models)

    script = __import__("train%d" % args.size)
    check_models()
    script.torch = proxy

    cfg = argparse.Namespace(path="~", dataset="vg", vg_dir="~/vg", batch_size=args.batch, niter=args.niter,
                             image_size=args.size, object_size=args.size // 2, embedding_dim=64, z_dim=64,
                             learning_rate=2e-4, resi_num=6, clstm_layers=3, lambda_img_adv=1.0, lambda_obj_adv=1.0,
                             lambda_obj_cls=1.0, lambda_z_rec=8.0, lambda_img_rec=1.0, lambda_kl=0.01, lambda_att_cls=2.0,
                             resume_iter="l", log_step=1, tensorboard_step=100, save_step=args.niter,
                             use_tensorboard=False, exp_name="harness")
    _, model_dir, _, _ = script.prepare_dir(cfg.exp_name)
    for key, appendix in (("G", "netG"), ("D_img", "netD_image"), ("D_obj", "netD_object"), ("D_att", "netD_attribute")):
        torch.save({k: v.clone() for k, v in states[key].items()}, os.path.join(model_dir, "iter-0_%s.pkl" % appendix))

    torch.manual_seed(0)
    random.seed(0)
    script.main(cfg)
    print("HARNESS_DONE model_dir=%s" % os.path.abspath(model_dir))


def run_test_script(args, torch, proxy, states, written, check_models):
    """test64.py:75-262 — generator in eval mode, attribute estimation, generation, attribute edit, second generation,
    attribute-classifier precision / recall.  The script parses sys.argv and runs main() at import time (test64.py:265),
    so the device proxy has to be what ITS `import torch` binds: the unmodified source is compiled and executed in a module
    namespace whose `__import__` hands out the proxy for the top-level torch package (nothing else sees the proxy)."""
    assert args.size == 64
    os.makedirs("data", exist_ok=True)
    if not os.path.exists("data/vocab.json"):
        os.symlink(os.path.join(REF, "data", "vocab.json"), "data/vocab.json")       # test64.py:82 opens it relative to cwd
    exp = "est_change_att_vg_bs%de64z64clstm3li1.0lo1.0lc1.0lz8.0lc1.0lk0.01" % args.batch    # test64.py:308-318
    g_dir = os.path.join("~", "checkpoints", "all", "models", exp)
    d_dir = os.path.join("~", "models", "trained_models")                             # test64.py:103
    for d, key, appendix in ((g_dir, "G", "netG"), (d_dir, "D_att", "netD_attribute")):
        os.makedirs(d, exist_ok=True)
        torch.save({k: v.clone() for k, v in states[key].items()}, os.path.join(d, "iter-0_%s.pkl" % appendix))
    sys.argv = ["test64.py", "--batch_size", str(args.batch), "--resume_iter", "l"]
    torch.manual_seed(0)
    random.seed(0)
    import builtins
    real_import = builtins.__import__

    def importer(name, globals=None, locals=None, fromlist=(), level=0):
        mod = real_import(name, globals, locals, fromlist, level)
        return proxy if (mod is torch) else mod           # `import torch` / `import torch.x.y as z` bind through the proxy

    ns_builtins = dict(vars(builtins))
    ns_builtins["__import__"] = importer
    path = os.path.join(REF, "test64.py")
    code = compile(open(path).read(), path, "exec")       # the unmodified source, executed as module `test64`
    exec(code, {"__name__": "test64", "__file__": path, "__builtins__": ns_builtins})
    check_models()
    torch.save(written, "written_images.pt")
    print("HARNESS_DONE images=%d" % len(written))


if __name__ == "__main__":
    main()
