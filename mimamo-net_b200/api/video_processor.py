"""Drop-in for api/video_processor.py: a thin subprocess wrapper around OpenFace's
FeatureExtraction binary.  No GPU work happens here; the class is kept so `Tester` wires up the
same way, and synthetic runs bypass it."""
import os
import subprocess


class Video_Processor(object):
    def __init__(self, size=112, nomask=True, grey=False, quiet=True,
                 tracked_vid=False, noface_save=False,
                 OpenFace_exe='OpenFace/build/bin/FeatureExtraction'):
        self.size = size
        self.nomask = nomask
        self.grey = grey
        self.quiet = quiet
        self.tracked_vid = tracked_vid
        self.noface_save = noface_save
        self.OpenFace_exe = OpenFace_exe
        if not isinstance(self.OpenFace_exe, str) or not os.path.exists(self.OpenFace_exe):
            raise ValueError("OpenFace_exe has to be string object and needs to exist.")
        self.OpenFace_exe = os.path.abspath(self.OpenFace_exe)

    def command(self, input_video, output_dir):
        """The FeatureExtraction argv the reference assembles (:69-83)."""
        argv = [self.OpenFace_exe, '-fdir' if os.path.isdir(input_video) else '-f', input_video,
                '-out_dir', output_dir, '-simsize', str(self.size),
                '-2Dfp', '-3Dfp', '-pdmparams', '-pose', '-aus', '-gaze', '-simalign']
        flags = [(not self.noface_save, '-nobadaligned'), (self.tracked_vid, '-tracked'),
                 (self.nomask, '-nomask'), (self.grey, '-g'), (self.quiet, '-q')]
        return argv + [f for on, f in flags if on]

    def process(self, input_video, output_dir=None):
        if not isinstance(input_video, str) or not os.path.exists(input_video):
            raise ValueError("input video has to be string object and needs to exist.")
        if os.path.isdir(input_video):
            assert len(os.listdir(input_video)) > 0, "Input sequence directory {} cannot be empty".format(input_video)
        input_video = os.path.abspath(input_video)
        if output_dir is None:
            output_dir = os.path.join(os.path.dirname(input_video), os.path.basename(input_video).split('.')[0])
        if not isinstance(output_dir, str):
            raise ValueError("output_dir should be string object.")
        if os.path.exists(output_dir):
            print("output dir exists: {}. Video processing skipped.".format(output_dir))
            return
        os.makedirs(output_dir)
        subprocess.call(self.command(input_video, output_dir))
