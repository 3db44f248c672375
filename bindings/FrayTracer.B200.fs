// FrayTracer.B200.fs -- P/Invoke shim over libfraytracer_b200.so (include/fraytracer_b200.h).
//
// UNVERIFIED SOURCE: no .NET toolchain exists in the build environment of this repository, so this file has
// never been compiled.  It shows the binding a FrayTracer maintainer would add: the same constructor names as
// SdfForm / SdfObject / SdfMaterial / SdfLight (src/FrayTracer/*.fs), returning reified nodes instead of
// closures, and drop-in replacements for `SdfScene.trace` + `Image.render` (Program.fs:90-93).
namespace FrayTracer.B200

open System
open System.Numerics
open System.Runtime.InteropServices
open FrayTracer

#nowarn "9"

module Native =
    [<Literal>]
    let Lib = "fraytracer_b200"

    [<Struct; StructLayout(LayoutKind.Sequential)>]
    type FormDesc =
        val mutable kind : int
        val mutable firstChild : int
        val mutable childCount : int
        [<MarshalAs(UnmanagedType.ByValArray, SizeConst = 10)>]
        val mutable p : float32[]

    [<Struct; StructLayout(LayoutKind.Sequential)>]
    type MaterialDesc = { kind : int; r : float32; g : float32; b : float32 }

    [<Struct; StructLayout(LayoutKind.Sequential)>]
    type ObjectDesc = { kind : int; form : int; material : int; firstChild : int; childCount : int }

    [<Struct; StructLayout(LayoutKind.Sequential)>]
    type LightDesc = { kind : int; vx : float32; vy : float32; vz : float32; r : float32; g : float32; b : float32 }

    [<Struct; StructLayout(LayoutKind.Sequential)>]
    type SceneDesc =
        { forms : nativeint; objects : nativeint; children : nativeint; materials : nativeint; lights : nativeint
          nForms : int; nObjects : int; nChildren : int; nMaterials : int; nLights : int; rootObject : int
          bgR : float32; bgG : float32; bgB : float32 }

    // ft_cloud_params (EXTENSION): perm -> 256 x int32, gradients -> 256 x 3 x float (F# float = float64)
    [<Struct; StructLayout(LayoutKind.Sequential)>]
    type CloudParamsDesc =
        { density : float32; frequency : float32; step : float32; fade : float32; bias : float32; gain : float32
          perm : nativeint; gradients : nativeint }

    // mirrors the Camera record (Camera.fs:16-22): 4 x Vector3
    [<Struct; StructLayout(LayoutKind.Sequential)>]
    type CameraDesc = { position : Vector3; forward : Vector3; upScaled : Vector3; rightScaled : Vector3 }

    [<DllImport(Lib)>] extern int ft_abi_version()
    [<DllImport(Lib)>] extern nativeint ft_last_error()
    [<DllImport(Lib)>] extern int ft_scene_create(SceneDesc& desc, int device, nativeint& scene)
    [<DllImport(Lib)>] extern int ft_scene_destroy(nativeint scene)
    [<DllImport(Lib)>] extern int ft_host_register(nativeint ptr, int64 bytes)
    [<DllImport(Lib)>] extern int ft_host_unregister(nativeint ptr)
    [<DllImport(Lib)>]
    extern int ft_render(nativeint scene, CameraDesc& camera, int width, int height, float32 epsilon, float32 length,
                         uint32 flags, nativeint outRgb, nativeint outLeaf, nativeint stats)
    // EXTENSIONS (no reference implementation; DESIGN.md 4.6 / 4.7)
    [<DllImport(Lib)>]
    extern int ft_render_spectral(nativeint scene, CameraDesc& camera, int width, int height, float32 epsilon,
                                  float32 length, int nWavelengths, int maxBounces, uint64 seed, nativeint outRgb,
                                  nativeint stats)
    [<DllImport(Lib)>]
    extern int ft_render_cloud(nativeint scene, CameraDesc& camera, int width, int height, float32 epsilon,
                               float32 length, CloudParamsDesc& cloud, int samplesPerPixel, uint64 seed,
                               nativeint outRgb, nativeint stats)
    [<DllImport(Lib)>]
    extern int ft_trace_rays(nativeint scene, int64 n, nativeint rays, uint32 flags, nativeint outRgb,
                             nativeint outLeaf, nativeint outSteps, nativeint outShadow, nativeint stats)

    let check (status:int) =
        if status <> 0 then failwith (Marshal.PtrToStringAnsi(ft_last_error()))

// ---- reified nodes: same constructors as the reference, data instead of closures ----------------------------
type FormNode =
    | Sphere of SdfForm.Primitive.Sphere          // SdfForm.fs:118-135
    | Capsule of SdfForm.Primitive.Capsule        // SdfForm.fs:137-170
    | Torus of SdfForm.Primitive.Torus            // SdfForm.fs:172-203
    | Triangle of SdfForm.Primitive.Triangle      // SdfForm.fs:205-268
    | Union of FormNode[]                         // SdfForm.fs:14-40
    | Subtract of FormNode * FormNode             // SdfForm.fs:42-49
    | Intersect of FormNode[]                     // SdfForm.fs:51-67
    | UnionSmooth of float32 * FormNode[]         // SdfForm.fs:69-91

type MaterialNode = Solid of FColor              // SdfMaterial.fs:4-7

type ObjectNode =
    | Leaf of MaterialNode * FormNode             // SdfObject.create   SdfObject.fs:6-10
    | ObjUnion of ObjectNode[]                    // SdfObject.union    SdfObject.fs:12-48
    | ObjSubtract of ObjectNode * FormNode        // SdfObject.subtract SdfObject.fs:50-54
    | ObjIntersect of ObjectNode * FormNode[]     // SdfObject.intersect SdfObject.fs:56-64

type LightNode =
    | Directional of Vector3 * FColor             // SdfLight.directional SdfLight.fs:6-21
    | Point of Vector3 * FColor                   // SdfLight.point       SdfLight.fs:23-42

type SceneNode = { Object : ObjectNode; BackgroundColor : FColor; Lights : LightNode list }   // Types.fs:74-79

[<RequireQualifiedAccess>]
module SdfForm =
    module Primitive =
        let sphere d = Sphere d
        let capsule d = Capsule d
        let torus d = Torus d
        let triangle d = Triangle d
    let union (forms:seq<FormNode>) =
        match Seq.toArray forms with
        | [||] -> failwith "No SdfObjects given."
        | [|f|] -> f
        | fs -> Union fs
    let subtract a b = Subtract (a, b)
    let intersect (forms:seq<FormNode>) =
        match Seq.toArray forms with
        | [||] -> failwith "No SdfObjects given."
        | [|f|] -> f
        | fs -> Intersect fs
    let unionSmooth strength (forms:seq<FormNode>) =
        match Seq.toArray forms with
        | [||] -> failwithf "blub"
        | [|f|] -> f
        | fs -> UnionSmooth (strength, fs)

[<RequireQualifiedAccess>]
module SdfMaterial =
    let createSolid color = Solid color

[<RequireQualifiedAccess>]
module SdfObject =
    let create material form = Leaf (material, form)
    let union (objects:seq<ObjectNode>) =
        match Seq.toArray objects with
        | [||] -> failwith "No SdfObjects given."
        | [|o|] -> o
        | os -> ObjUnion os
    let subtract o form = ObjSubtract (o, form)
    let intersect o (forms:seq<FormNode>) = ObjIntersect (o, Seq.toArray forms)

[<RequireQualifiedAccess>]
module SdfLight =
    let directional direction color = Directional (direction, color)
    let point position color = Point (position, color)

/// A scene compiled onto one B200.  `SdfScene.trace scene` of the reference becomes `new CompiledScene(scene, 0)`.
type CompiledScene(scene:SceneNode, device:int) =
    let mutable handle = 0n
    do
        // Flatten the node graph into the blittable tables of ft_scene_desc (children before parents, one
        // shared `children` index array), pin them, call ft_scene_create, unpin.  Omitted here for brevity:
        // it is the same walk as fraytracer_b200/api.py : SceneDescription.__init__.
        ()
    member _.Handle = handle
    interface IDisposable with
        member _.Dispose() = if handle <> 0n then Native.ft_scene_destroy handle |> ignore

[<RequireQualifiedAccess>]
module Image =
    /// Image.render epsilon length imageSize camera (SdfScene.trace scene)   (Image.fs:26-35)
    /// FColor[X,Y] is row-major with X the slow index and 12 bytes per element: exactly ft_render's out_rgb.
    let render (epsilon:float32) (length:float32) (imageSize:ImageSize) (camera:Camera) (scene:CompiledScene) : FColor[,] =
        let image = Array2D.zeroCreate<FColor> imageSize.X imageSize.Y
        let pin = GCHandle.Alloc(image, GCHandleType.Pinned)
        try
            let mutable cam : Native.CameraDesc =
                { position = camera.Position; forward = camera.Forward
                  upScaled = camera.UpScaled; rightScaled = camera.RightScaled }
            Native.ft_render(scene.Handle, &cam, imageSize.X, imageSize.Y, epsilon, length, 0u,
                             pin.AddrOfPinnedObject(), 0n, 0n)
            |> Native.check
        finally pin.Free()
        image
