"""Drop-in for the reference's ui/Sampling.py: the CLI-flavoured twin of ui/import_PC.py
(prints instead of callbacks, swallows every exception, ui/Sampling.py:21-80)."""
import os

from . import import_PC as _ipc


def process_chunk(points_chunk, las, voxel_size):
    """ui/Sampling.py:10-18 — `las` is accepted and ignored, as in the reference."""
    return _ipc.process_chunk(points_chunk, voxel_size)


def voxel_downsample_open3d(input_path, output_path, voxel_size, chunk_size=1000000):
    try:
        os.makedirs(os.path.dirname(output_path), exist_ok=True)
        print("正在读取输入文件...")
        print(f"开始分块处理（每块 {chunk_size} 个点）...")
        from .. import las as _las
        total_points = _las.read_header(input_path).point_count
        count = _ipc._downsample_file(input_path, output_path, voxel_size, chunk_size)
        print("合并处理结果...")
        print("正在写入输出文件...")
        print(f"\n成功生成下采样文件: {output_path}")
        print(f"原始点数: {total_points} → 下采样后点数: {count}")
    except Exception as e:
        print(f"\n处理过程中发生错误: {str(e)}")


if __name__ == "__main__":
    import sys
    if len(sys.argv) >= 3:
        voxel_downsample_open3d(sys.argv[1], sys.argv[2], float(sys.argv[3]) if len(sys.argv) > 3 else 0.1,
                                int(sys.argv[4]) if len(sys.argv) > 4 else 500000)
    else:
        print("usage: python -m pointcloudhookup_b200.ui.Sampling in.las out.las [voxel_size] [chunk_size]")
