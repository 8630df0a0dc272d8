"""Surface-compatible stand-in for the reference's ui/compress.py — the GIM container
(776-byte header + 7z payload) extractor/packer.  north_star lists it, but it touches no point data
(SURVEY.md §0): it is host file I/O with no kernel behind it, kept only so that
``from ui.compress import GIMExtractor`` keeps working.  py7zr is imported lazily (absent here)."""
import os
import shutil
import subprocess
import uuid
from io import BytesIO

GIM_HEADER_BYTES = 776


class GIMUtils:
    def generate_unique_filename(self):
        return f"{uuid.uuid4()}.7z"

    def get_filename(self, full_path):
        if not full_path.endswith(".gim"):
            raise ValueError("❌ 输入的文件路径不是以 .gim 结尾的")
        return os.path.basename(full_path)[: -len(".gim")]

    def ensure_folder_exists(self, folder_path):
        existed = os.path.exists(folder_path)
        os.makedirs(folder_path, exist_ok=True)
        print(f"📁 文件夹已存在: {folder_path}" if existed else f"✅ 已创建文件夹: {folder_path}")

    def read_file_to_parse(self, file_path):
        data = {}
        with open(file_path, "r", encoding="utf-8") as fh:
            for raw in fh:
                line = raw.strip()
                if line and "=" in line:
                    k, v = line.split("=", 1)
                    data[k.strip()] = v.strip()
        return data


utils = GIMUtils()


def _py7zr():
    try:
        import py7zr
        return py7zr
    except ImportError as e:  # pragma: no cover - depends on the host environment
        raise RuntimeError("py7zr is required for GIM archives and is not installed") from e


class GIMExtractor:
    def __init__(self, gim_file, output_folder="output"):
        self.gim_file = gim_file
        self.output_folder = output_folder
        self.gim_header = None

    def extract_embedded_7z(self):
        name = utils.get_filename(self.gim_file)
        print(f"🔄 正在解压文件：{self.gim_file}")
        with open(self.gim_file, "rb") as fh:
            self.gim_header = fh.read(GIM_HEADER_BYTES)
            payload = fh.read()
        utils.ensure_folder_exists(self.output_folder)
        target = os.path.join(self.output_folder, name)
        os.makedirs(target, exist_ok=True)
        with _py7zr().SevenZipFile(BytesIO(payload), mode="r") as archive:
            archive.extractall(path=target)
        print(f"✅ 解压完成，输出目录：{target}")
        return target

    def has_7z_cli(self):
        return shutil.which("7z") is not None

    def compress_with_7z_cli(self, source_folder, output_7z_path):
        subprocess.run(["7z", "a", "-mx=1", output_7z_path, source_folder], check=True)

    def compress_with_py7zr(self, source_folder):
        py7zr = _py7zr()
        buf = BytesIO()
        with py7zr.SevenZipFile(buf, "w", filters=[{"id": py7zr.FILTER_COPY}]) as archive:
            archive.writeall(source_folder, arcname="")
        return buf.getvalue()

    def build_custom_file(self, folder_to_compress, output_file, header_path=None):
        if header_path:
            with open(header_path, "rb") as hf:
                header = hf.read(GIM_HEADER_BYTES)
        else:
            header = self.gim_header
        if header is None or len(header) < GIM_HEADER_BYTES:
            raise ValueError("❌ Header 文件不足 776 字节")
        if self.has_7z_cli():
            print("🧰 使用系统 7z CLI 加速压缩")
            tmp = output_file + ".tmp.7z"
            self.compress_with_7z_cli(folder_to_compress, tmp)
            with open(tmp, "rb") as fh:
                payload = fh.read()
            os.remove(tmp)
        else:
            print("🐍 使用 py7zr 纯 Python 模式压缩（较慢）")
            payload = self.compress_with_py7zr(folder_to_compress)
        with open(output_file, "wb") as out:
            out.write(header)
            out.write(payload)
        print(f"✅ 封装完成: {output_file}")
