"""The Rust shim (rust/oar-ocr-b200) cannot be compiled here (no cargo / rustc in the image), so its FFI surface is
checked textually against include/oar_b200.h: every `pub fn oar_*` it declares exists in the header with the same
number of parameters, the status codes and model kinds agree, and the #[repr(C)] structs list the header's fields in
the header's order."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _strip_comments(s, line="//"):
    s = re.sub(r"/\*.*?\*/", "", s, flags=re.S)
    return re.sub(re.escape(line) + r".*", "", s)


def _header():
    return _strip_comments(open(os.path.join(ROOT, "include", "oar_b200.h")).read())


def _sys():
    return _strip_comments(open(os.path.join(ROOT, "rust", "oar-ocr-b200", "src", "sys.rs")).read())


def _c_functions(h):
    out = {}
    for m in re.finditer(r"\b(?:int32_t|int64_t|void|const char\*)\s+(oar_\w+)\s*\(([^;{]*?)\)\s*;", h, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    return out


def _rust_functions(r):
    out = {}
    for m in re.finditer(r"pub fn (oar_\w+)\s*\((.*?)\)\s*(?:->\s*[\w\*: ]+)?;", r, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if not args else len([a for a in args.split(",") if a.strip()])
    return out


def test_rust_declarations_match_the_header():
    c, r = _c_functions(_header()), _rust_functions(_sys())
    assert len(r) >= 15
    for name, n in r.items():
        assert name in c, f"{name} is not declared in include/oar_b200.h"
        assert c[name] == n, f"{name}: header has {c[name]} parameters, sys.rs {n}"
    for must in ("oar_model_load_onnx", "oar_det_run", "oar_rec_run", "oar_crop_rec_run", "oar_pipeline_run",
                 "oar_pipeline_run_multi"):
        assert must in r


def test_rust_constants_match_the_header():
    h, r = _header(), _sys()
    for name in ("OAR_OK", "OAR_E_INVALID", "OAR_E_NO_DEVICE", "OAR_E_CUDA", "OAR_E_MODEL", "OAR_E_CAPACITY",
                 "OAR_E_UNSUPPORTED", "OAR_KIND_DET", "OAR_KIND_REC", "OAR_KIND_CLS"):
        ch = re.search(rf"#define {name}\s+\(?(-?\d+)\)?", h)
        cr = re.search(rf"pub const {name}: i32 = (-?\d+);", r)
        assert ch and cr, name
        assert int(ch.group(1)) == int(cr.group(1)), name


def _c_struct_fields(h, name):
    body = re.search(r"typedef struct\s*\{([^}]*)\}\s*" + name + r"\s*;", h, flags=re.S).group(1)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        # "float ms_h2d, ms_det" / "int32_t* cols" / "oar_det_config det"
        names = decl.split(None, 1)[1] if " " in decl else decl
        for n in names.split(","):
            fields.append(re.sub(r"[\*\s]|\[.*\]", "", n.split()[-1] if " " in n.strip() else n))
    return fields


def _rust_struct_fields(r, name):
    body = re.search(r"pub struct " + name + r"\s*\{(.*?)\n\}", r, flags=re.S).group(1)
    return re.findall(r"pub (\w+):", body)


def test_rust_struct_layouts_match_the_header():
    h, r = _header(), _sys()
    for name in ("oar_det_config", "oar_pipeline_config", "oar_ocr_result"):
        assert _rust_struct_fields(r, name) == _c_struct_fields(h, name), name


def test_adapters_cover_the_trait_surface():
    src = open(os.path.join(ROOT, "rust", "oar-ocr-b200", "src", "adapters.rs")).read()
    for adapter in ("B200TextDetectionAdapter", "B200TextRecognitionAdapter"):
        assert f"impl ModelAdapter for {adapter}" in src
        assert f"impl AdapterBuilder for {adapter}Builder" in src
    for method in ("fn info(", "fn execute(", "fn supports_batching(", "fn recommended_batch_size(", "fn build(",
                   "fn with_config(", "fn adapter_type("):
        assert src.count(method) >= 2, method
