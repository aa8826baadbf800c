"""Device query helpers -- mirror of ``spim.process.cuda.{CUDADevice,CUDATools,NativeLibraryTools}``.

Reference (under /root/reference/src/main/java/spim/process/cuda/):
  CUDADevice.java:24-58, CUDATools.java:45-213 (queryCUDADetails), NativeLibraryTools.java:41-175.
The GUI dialogs of the reference are replaced by plain arguments; the queries go through the same
eight native symbols.
"""
from __future__ import annotations

import os
from typing import List, Optional

from . import native


class CUDADevice:
    """CUDADevice.java:24-58 -- id -1 denotes the CPU in the reference; never produced here."""

    def __init__(self, deviceId: int, deviceName: str, totalDeviceMemory: int, freeDeviceMemory: int,
                 majorVersion: int, minorVersion: int):
        self.deviceId = deviceId
        self.deviceName = deviceName
        self.totalDeviceMemory = totalDeviceMemory
        self.freeDeviceMemory = freeDeviceMemory
        self.majorVersion = majorVersion
        self.minorVersion = minorVersion

    def getDeviceName(self):
        return self.deviceName

    def getDeviceId(self):
        return self.deviceId

    def getTotalDeviceMemory(self):
        return self.totalDeviceMemory

    def getFreeDeviceMemory(self):
        return self.freeDeviceMemory

    def getMajorComputeVersion(self):
        return self.majorVersion

    def getMinorComputeVersion(self):
        return self.minorVersion

    def __lt__(self, other: "CUDADevice"):   # compareTo: more free memory first
        return self.freeDeviceMemory > other.freeDeviceMemory

    def __repr__(self):
        return (f"CUDADevice(id={self.deviceId}, name={self.deviceName!r}, total={self.totalDeviceMemory}, "
                f"free={self.freeDeviceMemory}, cc={self.majorVersion}.{self.minorVersion})")


class CUDATools:
    @staticmethod
    def queryCUDADetails(cuda: native.CUDAFourierConvolution, askForMultipleDevices: bool = False,
                         select: Optional[List[int]] = None) -> Optional[List[CUDADevice]]:
        """CUDATools.java:45-213: enumerate devices; returns None when the driver failed (-1) or no
        device exists.  ``select`` replaces the GenericDialog device choice."""
        n = cuda.getNumDevicesCUDA()
        if n == -1:
            print("Querying CUDA devices crashed, no devices available.")
            return None
        if n == 0:
            print("No CUDA devices detected.")
            return None
        devs = []
        for i in range(n):
            name = bytearray(256)
            cuda.getNameDeviceCUDA(i, name)
            nm = bytes(name).split(b"\0", 1)[0].decode(errors="replace")
            total = cuda.getMemDeviceCUDA(i)
            try:
                free = cuda.getFreeMemDeviceCUDA(i)
            except Exception:   # UnsatisfiedLinkError tolerated for old libraries (CUDATools.java:97-105)
                free = total
            devs.append(CUDADevice(i, nm, total, free, cuda.getCUDAcomputeCapabilityMajorVersion(i),
                                   cuda.getCUDAcomputeCapabilityMinorVersion(i)))
        if select is not None:
            return [devs[i] for i in select]
        if askForMultipleDevices:
            return devs
        return [sorted(devs)[0]]


class NativeLibraryTools:
    EXTENSIONS = (".so", ".lib", ".dll", ".dylib")   # NativeLibraryTools.java:154-175

    @staticmethod
    def loadNativeLibrary(potentialNames: Optional[List[str]] = None, directory: Optional[str] = None,
                          ) -> Optional[native.CUDAFourierConvolution]:
        """NativeLibraryTools.java:83-152: look in ``directory`` for a library, preferring names that contain
        one of ``potentialNames`` ('fftCUDA', 'FourierConvolutionCUDA'); returns None (after printing)
        when it cannot be loaded, like the reference's UnsatisfiedLinkError branch."""
        potentialNames = potentialNames or ["fftCUDA", "FourierConvolutionCUDA"]
        directory = directory or os.path.dirname(native.default_library_path())
        try:
            files = [f for f in sorted(os.listdir(directory)) if f.endswith(NativeLibraryTools.EXTENSIONS)]
        except OSError:
            files = []
        pick = None
        for f in files:
            if any(p.lower() in f.lower() for p in potentialNames):
                pick = f
                break
        if pick is None and files:
            pick = files[0]
        if pick is None:
            print(f"No native library found in '{directory}'.")
            return None
        try:
            return native.CUDAFourierConvolution(os.path.join(directory, pick))
        except OSError as e:
            print(f"Cannot load native library '{pick}': {e}")
            return None
