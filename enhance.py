"""Audio enhancement script - the reference's CLI (reference enhance.py:1-169) on the B200-native path.

Usage (same flags as the reference):
    # single file
    python enhance.py --checkpoint checkpoints/best_model.pth --input noisy.wav --output enhanced.wav
    # directory
    python enhance.py --checkpoint checkpoints/best_model.pth --input-dir noisy_audios/ --output-dir enhanced_audios/

Differences: the device is a CUDA (sm_100) GPU - there is no CPU path, ``--device cpu`` is an error; directory mode runs
the files in mixed-length batches (``--batch-size``, default 64) instead of one by one; ``--precision`` selects the
operand mode (fp16 default, fp32 accuracy mode).
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main(argv=None):
    parser = argparse.ArgumentParser(description="Enhance noisy audio using Hybrid Vision Transformer")
    parser.add_argument("--checkpoint", type=str, required=True, help="Path to model checkpoint")
    parser.add_argument("--config-dir", type=str, default=os.path.join(ROOT, "config"),
                        help="Directory containing configuration files")
    parser.add_argument("--input", type=str, default=None, help="Path to input noisy audio file")
    parser.add_argument("--output", type=str, default=None, help="Path to save enhanced audio file")
    parser.add_argument("--input-dir", type=str, default=None, help="Directory containing noisy audio files")
    parser.add_argument("--output-dir", type=str, default=None, help="Directory to save enhanced audio files")
    parser.add_argument("--extension", type=str, default=".wav", help="Audio file extension to process (directory mode)")
    parser.add_argument("--device", type=str, default="cuda", help="CUDA device to use for inference (cuda, cuda:1, ...)")
    parser.add_argument("--no-normalize", action="store_true", help="Disable audio normalization")
    parser.add_argument("--precision", type=str, default=None, choices=["fp16", "fp32", "bf16"],
                        help="operand mode of the CUDA plan (default: model config / fp16)")
    parser.add_argument("--batch-size", type=int, default=64, help="clips per batch in directory mode")
    args = parser.parse_args(argv)

    single_file_mode = args.input is not None and args.output is not None
    directory_mode = args.input_dir is not None and args.output_dir is not None
    if not single_file_mode and not directory_mode:
        parser.error("Must specify either:\n  --input and --output for single file mode, or\n"
                     "  --input-dir and --output-dir for directory mode")
    if single_file_mode and directory_mode:
        parser.error("Cannot use both single file and directory mode simultaneously")
    if not args.device.startswith("cuda"):
        parser.error("this build runs on a CUDA (sm_100) device only; there is no CPU path")

    import hvit_b200  # noqa: F401
    from hvit_b200.models import create_hybrid_vit
    from hvit_b200.inference import AudioEnhancer
    from hvit_b200.utils import load_all_configs, load_model_weights

    print("Loading configuration...")
    try:
        config = load_all_configs(args.config_dir)
    except Exception:  # noqa: BLE001  (the reference falls back to defaults the same way)
        print("Warning: Could not load config files. Using defaults.")
        config = {}

    print("\nCreating model...")
    model = create_hybrid_vit(config)
    if args.precision is not None:
        model.precision = args.precision
    print(f"Loading checkpoint from {args.checkpoint}")
    model = load_model_weights(args.checkpoint, model, device=args.device, strict=True)
    print(f"Model loaded successfully on {args.device}")

    print("\nInitializing audio enhancer...")
    audio_config = config.get("audio", {})
    enhancer = AudioEnhancer(model=model, device=args.device, sample_rate=audio_config.get("sample_rate", 16000),
                             n_fft=audio_config.get("n_fft", 512), hop_length=audio_config.get("hop_length", 128),
                             win_length=audio_config.get("win_length", 512))
    normalize = not args.no_normalize
    if single_file_mode:
        print("\nEnhancing single file...")
        print(f"Input: {args.input}")
        print(f"Output: {args.output}")
        enhancer.enhance_file(input_path=args.input, output_path=args.output, normalize=normalize)
        print("\nEnhancement complete!")
    else:
        print("\nEnhancing directory...")
        print(f"Input directory: {args.input_dir}")
        print(f"Output directory: {args.output_dir}")
        print(f"File extension: {args.extension}")
        enhancer.enhance_directory(input_dir=args.input_dir, output_dir=args.output_dir, extension=args.extension,
                                   normalize=normalize, batch_size=args.batch_size)
        print("\nAll files enhanced successfully!")
    return 0


if __name__ == "__main__":
    sys.exit(main())
