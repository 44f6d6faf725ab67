"""ChalametPIRError -- mirrors chalametpir_common/src/error.rs:8-50 for the variants reachable on the server and client paths."""
from __future__ import annotations

from ._lib import lib

# status code -> reference variant name (include/chalamet_b200.h chpir_status)
VARIANTS = {
    1: "InvalidMatrixDimension",
    2: "IncompatibleDimensionForMatrixMultiplication",
    3: "IncompatibleDimensionForRowVectorTransposedMatrixMultiplication",
    4: "FailedToDeserializeMatrixFromBytes",
    5: "EmptyKVDatabase",
    6: "ExhaustedAllAttemptsToBuild3WiseXorFilter",
    7: "ExhaustedAllAttemptsToBuild4WiseXorFilter",
    8: "RowNotDecodable",
    9: "DecodedRowNotPrependedWithDigestOfKey",
    10: "FailedToDeserializeFilterFromBytes",
    11: "KVDatabaseSizeTooLarge",
    12: "InvalidHintMatrix",
    13: "ArithmeticOverflowAddingQueryIndicator",
    14: "UnsupportedArityForBinaryFuseFilter",
    15: "InvalidResponseVector",
    16: "ImpossibleEncodedDBMatrixElementBitLength",
    17: "PendingQueryExistsForKey",
    18: "PendingQueryDoesNotExistForKey",
    50: "InvalidArgument",
    51: "BufferTooSmall",
    52: "IoFailed",
    53: "InvalidSavedServer",
    100: "CudaDeviceNotFound",
    101: "CudaAllocationFailed",
    102: "CudaTransferFailed",
    103: "CudaKernelLaunchFailed",
    104: "CudaKernelExecutionFailed",
    105: "CudaUnsupportedDevice",
    106: "CudaPeerAccessUnavailable",
    107: "NcclFailed",
    110: "HostAllocationFailed",
}


class ChalametPIRError(Exception):
    def __init__(self, code: int):
        self.code = int(code)
        self.variant = VARIANTS.get(self.code, lib.chpir_strerror(self.code).decode())
        detail = ""
        if self.code >= 100:
            detail = lib.chpir_last_cuda_error().decode()
        super().__init__(self.variant + (f" ({detail})" if detail else ""))


def check(rc: int) -> None:
    if rc != 0:
        raise ChalametPIRError(rc)
