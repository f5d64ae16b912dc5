// swift-tools-version:5.9
// SwiftPM host for the B200 build of MFAFFI: the reference's package (Package.swift:13-19 there) builds the C shim
// `Sources/MFAFFI` plus the `@_cdecl` Swift bridge `Sources/MFABridge`; here the Swift layer shrinks to a system-library
// target that re-exports libMFAFFI.so (built by universal-metal-flash-attention_b200/build.py) and a thin forwarding module
// for Swift callers.  Not compiled in this repository's CI image (no Swift toolchain); the C symbols are what is tested.
//
//   swift build -Xcc -I../include -Xlinker -L../universal-metal-flash-attention_b200/lib
import PackageDescription

let package = Package(
    name: "UniversalMetalFlashAttentionB200",
    products: [
        .library(name: "MFABridge", targets: ["MFABridge"]),
    ],
    targets: [
        .systemLibrary(name: "CMFACuda", path: "Sources/CMFACuda"),
        .target(name: "MFABridge", dependencies: ["CMFACuda"], path: "Sources/MFABridge"),
    ]
)
